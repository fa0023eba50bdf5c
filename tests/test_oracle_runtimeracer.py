"""CPU: the numpy restatement of the runtimeracer topology (oracle/runtimeracer_oracle.py, SURVEY.md section 8(f) row 1 --
groundwork, no product path yet) pinned to vectors minted from the unmodified reference (oracle/make_golden_rr.py)."""
import os

import numpy as np
import pytest

from oracle import runtimeracer_oracle as rr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("name,mode", [("rr_raw9.npz", "RAW"), ("rr_mol.npz", "MOL")])
def test_runtimeracer_oracle_matches_reference(name, mode):
    g = np.load(os.path.join(GOLD, name))
    sd = rr.make_state_dict_rr(seed=int(g["wseed"]), bits=9, mode=mode)
    tg, ov = int(g["target"]), int(g["overlap"])
    # teacher-forced on the reference's own samples: per-step logits (north_star: <= 1e-3 relative in fp32)
    o = rr.generate_rr(g["mel"], sd, int(g["seed"]), mode=mode, batched=True, target=tg, overlap=ov, forced=g["samples"], max_steps=48)
    assert o["logits"].shape == g["logits"].shape
    assert _rel(o["logits"], g["logits"]) < 1e-4
    # free running under the same injected noise: identical draws, same float64 wav
    f = rr.generate_rr(g["mel"], sd, int(g["seed"]), mode=mode, batched=True, target=tg, overlap=ov)
    if mode == "RAW":
        assert float((f["samples"][:, :-1] == g["samples"][:, :-1]).mean()) >= 0.999
    else:
        assert float((np.abs(f["samples"][:, :-1] - g["samples"][:, :-1]) < 1e-4).mean()) >= 0.999
    assert f["wav"].shape == g["wav"].shape and f["wav"].dtype == np.float64
    assert float(np.abs(f["wav"] - g["wav"]).max()) < 1e-3


def test_runtimeracer_teacher_forced_forward():
    """forward() of the reference (runtimeracer_version.py:136-196, nn.GRU over the whole sequence) equals the step-wise loop."""
    g = np.load(os.path.join(GOLD, "rr_raw9.npz"))
    sd = rr.make_state_dict_rr(seed=int(g["wseed"]), bits=9, mode="RAW")
    n = g["tf_logits"].shape[0]
    forced = np.zeros((1, n), np.float32)
    forced[0, :n - 1] = g["tf_x"][0, 1:n]                       # forward() is fed x[t]; the loop feeds back the sample of step t-1
    o = rr.generate_rr(g["mel"], sd, 0, mode="RAW", batched=False, forced=forced, max_steps=n)
    assert _rel(o["logits"][0], g["tf_logits"]) < 1e-4
