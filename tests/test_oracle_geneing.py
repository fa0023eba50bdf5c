"""CPU: the numpy restatement of the geneing topology (oracle/geneing_oracle.py, SURVEY.md section 8(f) row 3 -- groundwork, no
product path yet) pinned to vectors minted from the unmodified reference (oracle/make_golden_gn.py)."""
import os

import numpy as np

from oracle import geneing_oracle as gn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_geneing_oracle_matches_reference():
    g = np.load(os.path.join(GOLD, "gn_bits9.npz"))
    sd = gn.make_state_dict_gn(seed=int(g["wseed"]), bits=9)
    mels, aux = gn.upsample_network_generic(g["mel"], sd)                  # hop 200 = 4 x 5 x 10, 64 aux channels
    assert mels.shape == (24 * 200, 80) and aux.shape == (24 * 200, 64)
    assert _rel(mels[::37], g["up_mels"]) < 1e-5 and _rel(aux[::37], g["up_aux"]) < 1e-5
    tg, ov = int(g["target"]), int(g["overlap"])
    o = gn.generate_gn(g["mel"], sd, int(g["seed"]), target=tg, overlap=ov, forced=g["samples"], max_steps=48)
    assert _rel(o["logits"], g["logits"]) < 1e-4
    f = gn.generate_gn(g["mel"], sd, int(g["seed"]), target=tg, overlap=ov)
    assert float((f["samples"][:, :-1] == g["samples"][:, :-1]).mean()) >= 0.999
    assert f["wav"].shape == g["wav"].shape and float(np.abs(f["wav"] - g["wav"]).max()) < 1e-3
