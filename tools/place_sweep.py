"""Step time of the role-specialised loop against the physical placement of its groups (WRNN_RS_PLACE / WRNN_RS_ROT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
mol, _ = make_model(seed=12, bits=9, mode="MOL")
raw, _ = make_model(seed=11, bits=9, mode="RAW")
def setenv(**kw):
    for k, v in kw.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
def run(model, T, tg, ov, steps):
    mel = norm_mel(T, 1)
    out = model.generate_debug(mel, True, tg, ov, seed=3, max_steps=steps, precision=1)
    t = dict(model.last_timings)
    return t["ms_loop"] * 1e3 / t["n_steps"]
cases = [("raw b137", raw, 4800, 6000, 1000), ("mol b213", mol, 4800, 3000, 1500), ("raw b19", raw, 800, 8000, 800)]
rots = [int(a) for a in sys.argv[1:]] or list(range(0, 148, 8))
for name, model, T, tg, ov in cases:
    setenv(WRNN_RS_PLACE=None, WRNN_RS_ROT=None)
    base = min(run(model, T, tg, ov, 1500) for _ in range(2))
    line = []
    for r in rots:
        setenv(WRNN_RS_PLACE=1, WRNN_RS_ROT=r)
        line.append("%d:%.2f" % (r, min(run(model, T, tg, ov, 1500) for _ in range(2))))
    print(name, "scheduler's placement %.2f | by smid, rot " % base, " ".join(line), flush=True)
