"""CPU: the libwavernn `.bin` writer + Eigen-free C++ comparator against the oracle (dense and pruned weights)."""
import numpy as np
import pytest

from oracle import libwavernn_io, philox, weights, wavernn_oracle as orc
from oracle.libwavernn_runner import PortVocoder


@pytest.mark.parametrize("prune", [None, 0.9])
def test_port_matches_oracle(tmp_path, prune):
    sd = weights.make_state_dict(seed=11, bits=9, mode="RAW")
    if prune:
        sd = weights.prune_state_dict(sd, z=prune)
        W = sd["rnn2.weight_ih_l0"]
        zero = (np.abs(W).reshape(W.shape[0], -1, 4).sum(2) == 0).mean()
        assert 0.85 < zero <= 0.9                      # ties at the threshold are kept (pruner.py:80-81)
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    T = 4
    mel = weights.synthetic_mel(T, seed=5) / np.float32(4.0)
    U = philox.raw_uniforms(9, T * 200, 1)[:, 0]
    got = PortVocoder(path, 1).mel_to_wav(0, mel, U)
    mels, aux = orc.upsample_network(mel, sd)
    h1 = np.zeros((1, 512), np.float32); h2 = h1.copy(); x = np.zeros((1, 1), np.float32)
    want = np.zeros(T * 200, np.float32)
    for i in range(T * 200):
        lg, h1, h2 = orc.step_logits(x, mels[i:i + 1], aux[i:i + 1], h1, h2, sd)
        k = orc.sample_raw(lg, U[i:i + 1])
        want[i] = orc.label_to_float(k, 512)[0]
        x = want[i].reshape(1, 1)
    # -ffast-math turns the label division into a reciprocal multiply (1 ulp): compare class indices
    to_idx = lambda v: np.rint((v.astype(np.float64) + 1.0) * 511 / 2.0).astype(np.int64)
    assert (to_idx(got) == to_idx(want)).mean() >= 0.99


def test_threaded_vocode_mel_shape(tmp_path):
    sd = weights.prune_state_dict(weights.make_state_dict(seed=11, bits=9, mode="RAW"), z=0.9)
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    mel = weights.synthetic_mel(30, seed=5) / np.float32(4.0)
    wav = PortVocoder(path, 2).vocode_mel(mel, 1200, 400, 512)
    assert wav.shape == (30 * 200,) and np.isfinite(wav).all()
