"""Epilogue timeline of fold sets 0 and 1 in the tensor-core loop (CTA 0): where do the epilogue warps wait, where do they work?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WRNN_TC_TRACE"] = "gpurun_out/tc_trace.txt"
os.makedirs("gpurun_out", exist_ok=True)
import numpy as np
from tests.util import make_model, norm_mel
tg, ov = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "MOL"
model, _ = make_model(seed=12, bits=9, mode=mode)
mel = norm_mel(4800, 1)
model.generate_debug(mel, True, tg, ov, max_steps=200, precision=1, want_logits=False)
tr = np.loadtxt("gpurun_out/tc_trace.txt")
med = np.median(tr[2:], axis=0)
names = ["A enter", "x arrived", "A done", "B enter", "B acc", "B done", "C enter", "C acc", "C done", "D enter", "D acc", "D done"]
ev = []
for s in range(2):
    for k, n in enumerate(names):
        v = med[32 + 16 * s + k]
        if v >= 0 and (v > 0 or (s == 0 and k == 0)):
            ev.append((v, "set %d %s" % (s, n)))
for k, n in [(12, "prod ctr H1"), (16, "prod tma B issued"), (13, "prod ctr H2"), (17, "prod tma C issued"), (14, "prod ctr F1"), (18, "prod tma D issued")]:
    if med[k] > 0:
        ev.append((med[k], "   " + n + " (set 0)"))
for ph in range(4):
    for st_ in range(4):
        b = 64 + (ph * 4 + st_) * 4
        for k, n in enumerate(["ctr seen", "last tile issued", "first tile landed", "MMAs issued"]):
            if med[b + k] > 0:
                ev.append((med[b + k], "      queue: %s%d %s" % ("BCDE"[ph], st_, n)))
for kb in range(8):
    for base, n in [(128, 'slot free'), (136, 'landed'), (144, 'mma+commit issued')]:
        if med[base + kb] > 0:
            ev.append((med[base + kb], '         B1 tile %d %s' % (kb, n)))
ev.sort()
print("sets=%s tg=%d ov=%d %s: median SM clocks since set 0's A enter" % (os.environ.get("WRNN_TC_SETS", "auto"), tg, ov, mode))
for v, n in ev:
    print("%8.0f %7.2f us  %s" % (v, v / 1965.0, n))
d = np.diff(np.loadtxt("gpurun_out/tc_trace.txt.abs")) if os.path.exists("gpurun_out/tc_trace.txt.abs") else None
if d is not None:
    print("step period: %.2f us" % (np.median(d) / 1965.0))
