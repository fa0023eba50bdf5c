"""One short run of the role-specialised loop at the cfg3ref shape (MOL, 60 s, 213 folds), for ncu / sanitizer / timing.
   python tools/rs_run.py [max_steps] [seconds] [target] [overlap]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 60
target = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
overlap = int(sys.argv[4]) if len(sys.argv) > 4 else 1500
mode = os.environ.get("RS_MODE", "MOL")
bits = int(os.environ.get("RS_BITS", "9"))
model, _ = make_model(seed=12, bits=bits, mode=mode)
mel = norm_mel(80 * seconds, 1)
for it in range(2):
    t0 = time.time()
    out = model.generate_debug(mel, True, target, overlap, want_logits=False, seed=3, precision=1, max_steps=steps)
    dt = time.time() - t0
    t = dict(model.last_timings)
    print("%s run %d: %d folds x %d steps, loop %.3f ms = %.2f us/step (wall %.2f s)" % (t["loop_kernel"] + " " + mode, it, out["samples"].shape[0], out["samples"].shape[1],
          t["ms_loop"], t["ms_loop"] * 1e3 / out["samples"].shape[1], dt), flush=True)
assert np.isfinite(out["samples"]).all()
