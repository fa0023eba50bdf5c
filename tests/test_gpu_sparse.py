"""Block-sparse cluster loop (pruned checkpoints, BASELINE config 4) against the oracle on the same pruned weights."""
import numpy as np
import pytest

from oracle import wavernn_oracle as orc
from tests.util import make_model, norm_mel

pytestmark = pytest.mark.gpu
SPARSE = 2


@pytest.fixture(scope="module")
def pruned():
    return make_model(seed=11, bits=9, mode="RAW", prune=0.9)


def test_sparsity_detected(pruned):
    model, _ = pruned
    assert 0.85 < model.sparsity <= 0.9


@pytest.mark.parametrize("batched,tg,ov", [(True, 600, 100), (False, 0, 0)])
def test_sparse_loop_vs_oracle(pruned, batched, tg, ov):
    model, sd = pruned
    mel = norm_mel(23, 77)
    out = model.generate_debug(mel, batched, tg, ov, want_logits=True, seed=1234, max_steps=250, precision=SPARSE)
    S = tg + 2 * ov if batched else 23 * 200
    forced = np.pad(out["samples"], ((0, 0), (0, S - 250)))
    _, tr = orc.generate(mel, sd, mode="RAW", batched=batched, target=tg, overlap=ov, seed=1234, forced_samples=forced,
                         return_trace=True, max_steps=250)
    err = float(np.abs(out["logits"] - tr["logits"]).max() / np.abs(tr["logits"]).max())
    mine = np.rint((out["samples"] + 1.0) * 511 / 2.0).astype(np.int64)
    assert err < 1e-3, err
    assert float((mine == tr["index"]).mean()) >= 0.999


def test_sparse_equals_dense_kernels_on_pruned_weights(pruned):
    """The pruned checkpoint through the dense fp32 loop (zeros multiplied, like PyTorch) and through the sparse loop."""
    model, _ = pruned
    mel = norm_mel(60, 5)
    a = model.generate_debug(mel, True, 500, 100, want_logits=True, seed=3, max_steps=120)
    forced = np.pad(a["samples"], ((0, 0), (0, 700 - 120)))
    b = model.generate_debug(mel, True, 500, 100, forced=forced, want_logits=True, seed=3, max_steps=120, precision=SPARSE)
    assert a["samples"].shape[0] > 9          # more folds than one cluster takes: several clusters
    assert float(np.abs(a["logits"] - b["logits"]).max() / np.abs(a["logits"]).max()) < 1e-4
    assert float((a["samples"] == b["samples"]).mean()) >= 0.999


def test_sparse_full_generate(pruned):
    model, _ = pruned
    mel = norm_mel(40, 6)
    model.precision = SPARSE
    model.seed = 4
    wav = model.generate(mel[None], True, 1000, 200, True, True)
    model.precision = 0
    ref = model.generate(mel[None], True, 1000, 200, True, True)
    assert wav.shape == ref.shape and np.isfinite(wav).all()
    assert np.mean(np.abs(wav - ref) < 1e-6) > 0.9


def test_dense_checkpoint_rejected():
    model, _ = make_model(seed=11, bits=9, mode="RAW")
    with pytest.raises(ValueError):
        model.generate_debug(norm_mel(23, 1), True, 600, 100, max_steps=10, precision=SPARSE)
