import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
model, _ = make_model(seed=11, bits=9, mode="RAW", prune=0.9)
for name, T, batched, tg, ov in (("cfg2-like 3 s unbatched", 240, False, 0, 0), ("cfg4 10 s b19", 800, True, 8000, 800)):
    for prec, pname in ((2, "sparse cluster loop"), (0, "dense f32 loop"), (1, "dense f16 tc loop")):
        mel = norm_mel(T, 1)
        model.precision = prec
        best = None
        for _ in range(2):
            t0 = time.perf_counter(); wav = model.generate(mel[None], batched, tg, ov, True, True); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        t = model.last_timings
        print("%-26s %-20s loop %.1f ms  %.2f us/step  folds %d steps %d  => %.1fx RT" % (
            name, pname, t["ms_loop"], t["ms_loop"] * 1e3 / t["n_steps"], t["n_folds"], t["n_steps"], len(wav) / 16000.0 / best), flush=True)
