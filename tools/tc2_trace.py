import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WRNN_TC_TRACE"] = "gpurun_out/tc2_trace.txt"
os.makedirs("gpurun_out", exist_ok=True)
import numpy as np
from tests.util import make_model, norm_mel
model, _ = make_model(seed=12, bits=9, mode="MOL")
mel = norm_mel(4800, 1)
model.generate_debug(mel, True, 3000, 1500, max_steps=200, precision=1)
tr = np.loadtxt("gpurun_out/tc2_trace.txt")
names = {0: "step start", 1: "A: h1 exchanged (my part)", 2: "acc T0 ready", 3: "B: h2 exchanged", 4: "acc T2 ready", 5: "C: f1 exchanged",
         6: "acc T3 ready", 7: "D: f2 exchanged", 8: "acc T4 ready", 9: "sampled",
         10: "mma: T0 begin wait", 11: "mma: T1 begin", 12: "mma: T2 begin wait", 13: "mma: T3 begin wait", 14: "mma: T4 begin wait",
         15: "mma: T0 act ready", 16: "mma: T1 go", 17: "mma: T2 act ready", 18: "mma: T3 act ready", 19: "mma: T4 act ready",
         20: "mma: T0 issued", 21: "mma: T1 issued", 22: "mma: T2 issued", 23: "mma: T3 issued", 24: "mma: T4 issued"}
med = np.median(tr[2:], axis=0)
for k in sorted([k for k in names if med[k] >= 0], key=lambda k: med[k]):
    print("%8.0f  %6.2f us  %s" % (med[k], med[k] / 1965.0, names[k]))
