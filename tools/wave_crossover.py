"""loop_rs in waves against loop_tc at 256 .. 1024 folds (MOL): us per step of one launch over all folds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
mol, _ = make_model(seed=12, bits=9, mode="MOL")
def run(T, tg, ov, steps=1200):
    out = mol.generate_debug(norm_mel(T, 1), True, tg, ov, seed=3, max_steps=steps, precision=1)
    t = dict(mol.last_timings)
    return t["ms_loop"] * 1e3 / t["n_steps"] / max(1, t["n_launches"]), t["n_folds"], t["loop_kernel"], t["n_launches"]
for tg, ov in [(3410, 341), (2200, 360), (2270, 227), (1705, 170), (1140, 114), (853, 85)]:
    line = []
    for rs in ("1", "0"):
        os.environ["WRNN_RS"] = rs
        us, nf, k, nl = min(run(4800, tg, ov) for _ in range(2))
        line.append("%s x%d %.2f us/step" % (k, nl, us))
    print("%d/%d folds %d: " % (tg, ov, nf) + " | ".join(line), flush=True)
