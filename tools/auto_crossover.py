"""Where the fp32 loop stops being the faster one (the facade's AUTO_F16_MIN_FOLDS): loop us/step of both loops vs fold count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
model, _ = make_model(seed=11, bits=9, mode="RAW")
for T in [int(a) for a in sys.argv[1:]] or [800, 1000, 1200, 1600]:
    mel = norm_mel(T, 1)
    row = []
    for prec in (0, 1):
        model.precision = prec
        best = None
        for _ in range(2):
            model.generate(mel[None], True, 8000, 800, True, True)
            t = dict(model.last_timings)
            if best is None or t["ms_loop"] < best["ms_loop"]:
                best = t
        row.append(best["ms_loop"] * 1e3 / best["n_steps"])
    print("T=%d folds %d  fp32 %.2f us/step  fp16 %.2f us/step" % (T, best["n_folds"], row[0], row[1]), flush=True)
