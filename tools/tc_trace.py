import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WRNN_TC_TRACE"] = "gpurun_out/tc_trace.txt"
os.makedirs("gpurun_out", exist_ok=True)
import numpy as np
from tests.util import make_model, norm_mel
mode = sys.argv[1] if len(sys.argv) > 1 else "MOL"
shape = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
model, _ = make_model(seed=12, bits=9, mode=mode)
model.precision = 1
if shape == "cfg3":
    mel = norm_mel(4800, 1)
    model.generate_debug(mel, True, 3000, 1500, max_steps=200, precision=1)
elif shape == "b512":
    mel = norm_mel(4800, 1)
    model.generate_debug(mel, True, 1705, 170, max_steps=200, precision=1)
else:
    mel = norm_mel(800, 1)
    model.generate_debug(mel, True, 8000, 800, max_steps=200, precision=1)
tr = np.loadtxt("gpurun_out/tc_trace.txt")
names = {0: "step start", 1: "x arrived", 2: "A published", 3: "accB ready", 4: "B published", 5: "accC ready", 6: "C published",
         7: "accD ready", 8: "D published", 9: "accE ready", 10: "sampled",
         12: "prod: ctr H1", 13: "prod: ctr H2", 14: "prod: ctr F1", 15: "prod: ctr F2",
         16: "prod: tma B issued", 17: "prod: tma C issued", 18: "prod: tma D issued", 19: "prod: tma E issued",
         20: "mma: B tile0 full", 21: "mma: B tile1 full", 22: "mma: B tile2 full", 23: "mma: B tile3 full",
         24: "mma: B tile0 wait begins", 25: "mma: B tile1 wait begins", 26: "mma: B tile2 wait begins", 27: "mma: B tile3 wait begins",
         31: "set 1: step start", 28: "B: tmem loaded", 29: "B: math+store done", 30: "B: after bar.sync"}
med = np.median(tr[2:], axis=0)
order = sorted([k for k in names if med[k] >= 0], key=lambda k: med[k])
print("median SM clocks since step start (steps 66..79), %s %s" % (mode, shape))
for k in order:
    print("%8.0f  %6.2f us  %s" % (med[k], med[k] / 1965.0, names[k]))
print("step period (clocks):", np.median(np.diff(tr[:, 0] + 0)) if False else "n/a")
