"""Ad-hoc timing of the named configs on one GPU (not the bench contract; see bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import make_model, norm_mel

res = {}
which = sys.argv[1:] or ["f32", "f16"]

def run(model, name, T, batched, tg, ov, prec, reps=2):
    mel = norm_mel(T, 1)
    model.precision = prec
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        wav = model.generate(mel[None], batched, tg, ov, True, True)
        dt = time.perf_counter() - t0
        t = dict(model.last_timings)
        t["wall_ms"] = dt * 1e3
        if best is None or dt < best["wall_ms"] / 1e3:
            best = t
    best["us_per_step"] = best["ms_loop"] * 1e3 / best["n_steps"]
    best["x_realtime_e2e"] = (len(wav) / 16000.0) / (best["wall_ms"] / 1e3)
    res[name] = best
    print("%-28s loop %.1f ms  %.2f us/step  cond %.2f post %.2f h2d %.2f d2h %.2f wall %.1f ms  folds %d steps %d  => %.1fx RT" % (
        name, best["ms_loop"], best["us_per_step"], best["ms_cond"], best["ms_post"], best["ms_h2d"], best["ms_d2h"],
        best["wall_ms"], best["n_folds"], best["n_steps"], best["x_realtime_e2e"]), flush=True)

raw, _ = make_model(seed=11, bits=9, mode="RAW")
mol, _ = make_model(seed=12, bits=9, mode="MOL")
print("floor", raw.barrier_floor(20000), flush=True)
for tag in which:
    prec = 0 if tag == "f32" else 1
    if prec == 0:
        run(raw, "f32 cfg2 unbatched 1s", 80, False, 0, 0, prec)
    run(raw, tag + " cfg1 raw9 10s b19", 800, True, 8000, 800, prec)
    run(raw, tag + " raw9 60s b137", 4800, True, 6000, 1000, prec, reps=1 if prec == 0 else 2)
    run(mol, tag + " cfg3 mol 60s b137", 4800, True, 6000, 1000, prec, reps=1 if prec == 0 else 2)
    run(mol, tag + " cfg3 mol 60s b213", 4800, True, 3000, 1500, prec, reps=1 if prec == 0 else 2)
    if prec == 1:
        run(mol, tag + " mol 60s b256 3410/341", 4800, True, 3410, 341, prec)
        run(mol, tag + " mol 60s b512 1705/170", 4800, True, 1705, 170, prec)
        run(raw, tag + " raw 60s b512 1705/170", 4800, True, 1705, 170, prec)
        run(mol, tag + " mol 60s b766 1140/114", 4800, True, 1140, 114, prec)
        run(raw, tag + " raw 60s b766 1140/114", 4800, True, 1140, 114, prec)
        run(mol, tag + " mol 60s b1024 853/85", 4800, True, 853, 85, prec)
        run(mol, tag + " mol 60s b896 975/97", 4800, True, 975, 97, prec)
        run(raw, tag + " raw 60s b1024 853/85", 4800, True, 853, 85, prec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)
