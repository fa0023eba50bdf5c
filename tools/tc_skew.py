"""Cross-CTA skew of the tensor-core loop: %globaltimer stamps of every unit-owning CTA at a few events of one step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WRNN_TC_TRACE"] = "gpurun_out/tc_trace.txt"
os.makedirs("gpurun_out", exist_ok=True)
import numpy as np
from tests.util import make_model, norm_mel
tg, ov = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "MOL"
model, _ = make_model(seed=12, bits=9, mode=mode)
model.generate_debug(norm_mel(4800, 1), True, tg, ov, max_steps=200, precision=1, want_logits=False)
x = np.loadtxt("gpurun_out/tc_trace.txt.skew")
names = ["x arrived", "A done", "B acc ready", "D done"]
t0 = x[x > 0].min()
print("tg=%d ov=%d %s pair=%s: us since the earliest stamp; per event: min / median / max over the CTAs that serve the set" % (
    tg, ov, mode, os.environ.get("WRNN_TC_PAIR", "auto")))
for g in range(2):
    for s in range(4):
        for k, n in enumerate(names):
            v = x[64 * g:64 * g + 64, 4 * s + k]
            v = v[v > 0]
            if v.size:
                v = (v - t0) / 1e3
                print("group %d local set %d %-12s n=%2d  min %7.2f  med %7.2f  max %7.2f  spread %5.2f" % (g, s, n, v.size, v.min(), np.median(v), v.max(), v.max() - v.min()))
