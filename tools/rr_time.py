"""Step time of the runtimeracer loop (csrc/loop_rr.cu): python tools/rr_time.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_rr import make_rr
from tests.util import norm_mel

model, _ = make_rr(5, 9, "RAW")
for name, batched, sec, tg, ov, steps in [("unbatched 1 fold", False, 2, 0, 0, 4000), ("10 s, 8000/800", True, 10, 8000, 800, 2000),
                                          ("60 s, 6000/1000", True, 60, 6000, 1000, 1000), ("60 s, 3000/1500", True, 60, 3000, 1500, 1000)]:
    mel = norm_mel(80 * sec, 1)
    for it in range(2):
        out = model.generate_debug(mel, batched, tg, ov, seed=3, max_steps=steps)
    t = dict(model.last_timings)
    F, S = out["samples"].shape
    waves = (F + 63) // 64
    print("%-18s: %3d folds (%d launches) x %d steps: loop %.2f ms = %.2f us per step and wave" % (name, F, t["n_launches"], S, t["ms_loop"], t["ms_loop"] * 1e3 / S / waves))
