"""Time one loop configuration (environment switches are read by the engine per call): python tools/one_time.py MODE T target overlap [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel

mode, T, tg, ov = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
model, _ = make_model(seed=11 if mode == "RAW" else 12, bits=9, mode=mode)
model.precision = 1
mel = norm_mel(T, 1)
best = None
for _ in range(reps):
    model.generate(mel[None], True, tg, ov, True, True)
    t = dict(model.last_timings)
    if best is None or t["ms_loop"] < best["ms_loop"]:
        best = t
tags = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("WRNN_"))
print("%-4s T=%d %d/%d [%s] loop %.1f ms %.2f us/step folds %d steps %d  expand %.2f ms" % (
    mode, T, tg, ov, tags, best["ms_loop"], best["ms_loop"] * 1e3 / best["n_steps"], best["n_folds"], best["n_steps"],
    best.get("ms_expand", -1)), flush=True)
