"""Off-path ingest delay of T1 (WRNN_RS_DELAY_NS) against the step time, inline conditioning."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
mol, _ = make_model(seed=12, bits=9, mode="MOL")
raw, _ = make_model(seed=11, bits=9, mode="RAW")
def run(model, T, tg, ov, steps=1500):
    out = model.generate_debug(norm_mel(T, 1), True, tg, ov, seed=3, max_steps=steps, precision=1)
    t = dict(model.last_timings)
    return t["ms_loop"] * 1e3 / t["n_steps"]
for name, model, T, tg, ov in [("mol b213", mol, 4800, 3000, 1500), ("mol b137", mol, 4800, 6000, 1000), ("raw b19", raw, 800, 8000, 800)]:
    line = []
    for d in [int(a) for a in sys.argv[1:]] or [0, 1000, 2000, 2500, 3000, 4000]:
        os.environ["WRNN_RS_DELAY_NS"] = str(d)
        line.append("%d:%.2f" % (d, min(run(model, T, tg, ov) for _ in range(3))))
    print(name, "delay ns -> us/step ", " ".join(line), flush=True)
