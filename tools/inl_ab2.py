"""Group-count / inline A/B at small fold counts (alternating, best of 3)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
mol, _ = make_model(seed=12, bits=9, mode="MOL")
raw, _ = make_model(seed=11, bits=9, mode="RAW")
def setenv(**kw):
    for k, v in kw.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
def run(model, T, tg, ov):
    mel = norm_mel(T, 1)
    model.precision = 1
    model.generate(mel[None], True, tg, ov, True, True)
    t = dict(model.last_timings)
    return t["ms_loop"] * 1e3 / t["n_steps"], t["n_folds"]
cases = [("raw b19", raw, 800, 8000, 800), ("raw b137", raw, 4800, 6000, 1000), ("mol b19", mol, 800, 8000, 800), ("mol b66", mol, 2400, 6500, 650)]
variants = [(0, 1, 1), (1, 1, 0), (1, 1, 1), (1, 2, 0), (1, 2, 1), (1, 3, 1)]
for name, model, T, tg, ov in cases:
    best = {}
    for rep in range(3):
        for inl, G, pad in variants:
            setenv(WRNN_RS_INLINE=inl, WRNN_RS_GROUPS=G, WRNN_RS_PAD=pad)
            us, nf = run(model, T, tg, ov)
            best[(inl, G, pad)] = min(best.get((inl, G, pad), 1e9), us)
    print(name, "folds", nf, "  ".join("inl=%d G=%d pad=%d: %.2f" % (i, g, q, best[(i, g, q)]) for i, g, q in variants), flush=True)
