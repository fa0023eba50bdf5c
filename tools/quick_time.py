"""Ad-hoc timing of the named configs on one GPU (not the bench contract; see bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import make_model, norm_mel

res = {}
model, _ = make_model(seed=11, bits=9, mode="RAW")
res["floor"] = model.barrier_floor(20000)
print("floor", res["floor"], flush=True)

def run(model, name, T, batched, tg, ov, reps=2):
    mel = norm_mel(T, 1)
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
    print(name, json.dumps(best), flush=True)

run(model, "cfg2_unbatched_1s", 80, False, 0, 0)
run(model, "cfg1_raw9_10s_b19", 800, True, 8000, 800)
run(model, "raw9_60s_b137", 4800, True, 6000, 1000, reps=1)
mol, _ = make_model(seed=12, bits=9, mode="MOL")
run(mol, "cfg3_mol_60s_b137", 4800, True, 6000, 1000, reps=1)
run(mol, "cfg3_mol_60s_b213", 4800, True, 3000, 1500, reps=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)
