"""RAW: T1's sleep before it polls the sample word (WRNN_RS_RAW_SLEEP, % of the previous step's wait) against the step time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
raw, _ = make_model(seed=11, bits=9, mode="RAW")
raw10, _ = make_model(seed=11, bits=10, mode="RAW")
def run(model, T, tg, ov, steps=2000):
    out = model.generate_debug(norm_mel(T, 1), True, tg, ov, seed=3, max_steps=steps, precision=1)
    t = dict(model.last_timings)
    return t["ms_loop"] * 1e3 / t["n_steps"]
for name, model, T, tg, ov in [("raw9 b19", raw, 800, 8000, 800), ("raw9 b137", raw, 4800, 6000, 1000), ("raw9 b213", raw, 4800, 3000, 1500), ("raw10 b137", raw10, 4800, 6000, 1000)]:
    line = []
    for pct in [int(a) for a in sys.argv[1:]] or [0, 20, 30, 40, 45, 0]:
        os.environ["WRNN_RS_RAW_SLEEP"] = str(pct)
        line.append("%d%%:%.2f" % (pct, min(run(model, T, tg, ov) for _ in range(3))))
    print(name, " ".join(line), flush=True)
