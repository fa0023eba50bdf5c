"""The runtimeracer topology on the GPU (csrc/loop_rr.cu + engine_rr.inc) against vectors minted
from the UNMODIFIED reference (tests/golden/rr_*.npz, oracle/make_golden_rr.py) and against the oracle
(oracle/runtimeracer_oracle.py).  Gates: teacher-forced logits within 1e-4 relative (fp32 loop; the two FC pairs without an
activation between them are fused in float64, hence not bit-exact), >= 99.9 % identical draws, float64 wav within 1e-3."""
import copy
import os

import numpy as np
import pytest

from oracle import runtimeracer_oracle as rr
from tests.util import norm_mel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def make_rr(seed, bits=9, mode="RAW", device=0):
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_runtimeracer)
    hp.bits, hp.mode = bits, mode
    sd = rr.make_state_dict_rr(seed=seed, bits=bits, mode=mode)
    model, _ = base.init_voc_model(base.MODEL_TYPE_RUNTIMERACER, device, override_hp_runtimeracer=hp)
    model.load_state_dict(sd)
    assert base.get_model_type(model) == base.MODEL_TYPE_RUNTIMERACER
    return model, sd


@pytest.mark.parametrize("name,mode", [("rr_raw9.npz", "RAW"), ("rr_mol.npz", "MOL")])
def test_runtimeracer_gpu_matches_reference_golden(name, mode):
    g = np.load(os.path.join(GOLD, name))
    model, sd = make_rr(int(g["wseed"]), 9, mode)
    tg, ov = int(g["target"]), int(g["overlap"])
    S = tg + 2 * ov
    ref = g["samples"]
    forced = np.zeros((ref.shape[0], S), np.float32)
    forced[:, :ref.shape[1]] = ref
    n = g["logits"].shape[1]
    # teacher-forced on the reference's own samples: per-step logits
    o = model.generate_debug(g["mel"], True, tg, ov, forced=forced, want_logits=True, seed=int(g["seed"]), max_steps=n)
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rr_kernel"
    err = _rel(o["logits"], g["logits"])
    # free running under the same noise: the draws and the float64 waveform
    model.seed = int(g["seed"])
    f = model.generate_debug(g["mel"], True, tg, ov, seed=int(g["seed"]))
    k = ref.shape[1] - 1
    if mode == "RAW":
        agree = float((f["samples"][:, :k] == ref[:, :k]).mean())
    else:
        agree = float((np.abs(f["samples"][:, :k] - ref[:, :k]) < 1e-4).mean())
    wav = model.generate(g["mel"][None], True, tg, ov, True, True, seed=int(g["seed"]))
    werr = float(np.abs(wav - g["wav"]).max())
    print("runtimeracer %s vs reference golden: logits rel err %.3e, draw agreement %.5f, wav max err %.3e" % (mode, err, agree, werr))
    assert err < 1e-4, err
    assert agree >= 0.999, agree
    assert wav.shape == g["wav"].shape and wav.dtype == np.float64
    assert werr < 1e-3, werr


def test_runtimeracer_gpu_waves_and_unbatched_vs_oracle():
    """More folds than one launch holds (83 folds = two waves of <= 64) and the unbatched path (one fold, every sample a step):
    first steps against the oracle, teacher-forced on the kernel's own samples."""
    model, sd = make_rr(5, 9, "RAW")
    for batched, T, tg, ov, steps in [(True, 330, 700, 100, 24), (False, 12, 0, 0, 64)]:
        mel = norm_mel(T, 4)
        o = model.generate_debug(mel, batched, tg, ov, want_logits=True, seed=7, max_steps=steps)
        F = o["samples"].shape[0]
        S = tg + 2 * ov if batched else T * 200
        forced = np.zeros((F, S), np.float32)
        forced[:, :steps] = o["samples"]
        t = rr.generate_rr(mel, sd, 7, mode="RAW", bits=9, batched=batched, target=tg, overlap=ov, forced=forced, max_steps=steps)
        err = _rel(o["logits"], t["logits"])
        agree = float((o["samples"] == t["samples"]).mean())
        print("runtimeracer %s, %d folds x %d steps vs oracle: logits rel err %.3e, draw agreement %.5f" % ("batched" if batched else "unbatched", F, steps, err, agree))
        assert (F > 64) == batched
        assert err < 1e-4 and agree >= 0.999, (err, agree)


def test_runtimeracer_facade():
    """The drop-in facade with the model type the reference hard-codes for its C++ path (vocoder/inference.py:43)."""
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import inference
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_runtimeracer)
    hp.bits = 9
    sd = rr.make_state_dict_rr(seed=3, bits=9, mode="RAW")
    inference.load_state(sd, base.MODEL_TYPE_RUNTIMERACER, override_hp_runtimeracer=hp)
    try:
        inference.set_seed(1)
        mel = norm_mel(40, 2) * 4.0
        w1 = inference.infer_waveform(mel, target=1000, overlap=200)
        inference.set_seed(1)
        w2 = inference.infer_waveform(mel, target=1000, overlap=200)
        assert w1.shape == (39 * 200,) and w1.dtype == np.float64 and np.isfinite(w1).all()
        assert np.array_equal(w1, w2)
    finally:
        inference.unload()
