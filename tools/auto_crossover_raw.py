"""fp32 loop vs the role-specialised loop for RAW (9-bit) by fold count: python tools/auto_crossover_raw.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
model, _ = make_model(seed=11, bits=9, mode="RAW")
for T in (8, 12, 17, 22, 32, 42, 62):
    mel = norm_mel(T, 3)
    res = []
    for prec in (0, 1):
        for _ in range(2):
            out = model.generate_debug(mel, True, 800, 200, seed=3, max_steps=1024, precision=prec)
        t = dict(model.last_timings)
        res.append((t["loop_kernel"], t["ms_loop"] * 1e3 / 1024))
    print("%2d folds: %s %.2f us/step | %s %.2f us/step" % (out["samples"].shape[0], res[0][0], res[0][1], res[1][0], res[1][1]), flush=True)
