"""Layout calibration of the role-specialised loop: the trials (WRNN_VERBOSE) and the step time with / without it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
def run(model, T, tg, ov):
    model.precision = 1
    model.generate(norm_mel(T, 1)[None], True, tg, ov, True, True)
    t = dict(model.last_timings)
    return t["ms_loop"] * 1e3 / t["n_steps"] / max(1, t["n_launches"])
for name, mode, seed, T, tg, ov in [("cfg1 raw9 b19", "RAW", 11, 800, 8000, 800), ("raw9 b137", "RAW", 11, 4800, 6000, 1000), ("raw9 b35 3000/1500", "RAW", 11, 800, 3000, 1500),
                                    ("mol b213", "MOL", 12, 4800, 3000, 1500), ("mol b35", "MOL", 12, 800, 3000, 1500), ("mol b68", "MOL", 12, 2400, 6500, 650)]:
    res = {}
    for cal in ("0", "1"):
        os.environ["WRNN_RS_CALIBRATE"] = cal
        os.environ["WRNN_VERBOSE"] = "1" if cal == "1" else ""
        if cal == "0": os.environ.pop("WRNN_VERBOSE")
        model, _ = make_model(seed=seed, bits=9, mode=mode)      # a fresh engine: calibrates once
        run(model, T, tg, ov)
        res[cal] = min(run(model, T, tg, ov) for _ in range(2))
    print("%s: default layout %.2f us/step, calibrated %.2f" % (name, res["0"], res["1"]), flush=True)
