"""Timeline of the role-specialised loop (WRNN_RS_TRACE): per role, the mean time of every event relative to the moment
the LAST T4 CTA of the group published f2(t-1) (= the start of step t's chain), over the traced steps.
   python tools/rs_trace.py [workload] [trace file]"""
import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3ref"
path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "rs_trace.txt")
if not os.path.exists(path) or os.environ.get("RS_TRACE_RUN"):
    env = dict(os.environ, WRNN_RS_TRACE=path)
    subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-extras"], env=env,
                   stdout=subprocess.DEVNULL, check=True)
rows = np.loadtxt(path, dtype=np.int64)
ncta = rows[:, 0].max() + 1
NC = int(os.environ.get("RS_CTAS", "48"))
ev = rows[:, 2:].reshape(ncta, 8, 48).astype(np.float64)
ev[ev == 0] = np.nan
names = {0: "T1 GRU1+fc3", 1: "T2 GRU2", 2: "T3 fc1", 3: "T4 fc2", 4: "T5 sampler"}
evn = ["step top", "in1 canaries", "in1 loaded", "in1 in TMEM", "MMA done", "published", "in2 canaries", "in2 loaded", "in2 in TMEM",
       "mma1 first kq", "mma1 last kq", "mma2 first kq", "mma2 last kq"]
def passes(cs, k):
    a, b = ev[cs, k, 13], ev[cs, k, 14]
    return " | passes in1 %s in2 %s" % (np.nanmax(a) if not np.all(np.isnan(a)) else "-", np.nanmax(b) if not np.all(np.isnan(b)) else "-")
def role_of(c):
    r = c % NC
    return 0 if r < 16 else (1 if r < 32 else (2 if r < 40 else (3 if r < 48 else 4)))
G = ncta // NC
for g in range(min(G, 1)):
    ctas = np.arange(g * NC, (g + 1) * NC)
    roles = np.array([role_of(c) for c in ctas])
    # chain origin of step k: last T4 publish of step k-1 (event 5)
    t4 = ctas[roles == 3]
    print("group %d" % g)
    for k in range(1, 8):
        origin = np.nanmax(ev[t4, k - 1, 5])
        line = "  step +%d (chain origin = last f2 publish): " % k
        for r in range(5 if NC > 48 else 4):
            cs = ctas[roles == r]
            line += "\n    %-12s" % names[r]
            for j in range(13):
                v = ev[cs, k, j] - origin
                if np.all(np.isnan(v)):
                    continue
                line += " | %s %.2f..%.2f" % (evn[j], np.nanmin(v) / 1e3, np.nanmax(v) / 1e3)
            line += passes(cs, k)
            cl = np.nanmean(ev[cs, 1:8, 32:44], axis=(0, 1))          # SM clocks of warp 0, mean over CTAs of the role and steps
            nm = {0: "wait MMA", 1: "MMA done", 2: "logits loaded", 3: "drawn", 4: "acc loaded", 5: "math done", 6: "published",
                  8: "canaries", 9: "loaded", 10: "checked", 11: "in TMEM"}
            order = [8, 9, 10, 11, 0, 1, 2, 3, 4, 5, 6]
            base = cl[8]
            line += "\n        warp-0 clocks since its canaries: " + " | ".join("%s %.0f" % (nm[j], cl[j] - base) for j in order if not np.isnan(cl[j]))
            line += "\n        K quarters in TMEM (warp 0):"
            for q in range(4):
                v = ev[cs, k, 16 + q] - origin
                line += " %d: %.2f..%.2f" % (q, np.nanmin(v) / 1e3, np.nanmax(v) / 1e3)
        print(line)
        if k >= 2:
            break
    step = np.nanmax(ev[t4, 7, 5]) - np.nanmax(ev[t4, 1, 5])
    print("  mean step time over 6 steps: %.2f us" % (step / 6e3))
