"""A/B of the RAW tensor-core loop: dedicated sampler CTAs (default) vs classes spread over the unit-owning CTAs
(WRNN_TC_RAWSAMP=0).  Ad-hoc timing, not the bench contract."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel

raw, _ = make_model(seed=11, bits=9, mode="RAW")
mol, _ = make_model(seed=12, bits=9, mode="MOL")
raw.precision = 1
mol.precision = 1


def run(model, name, T, tg, ov, reps=2):
    mel = norm_mel(T, 1)
    best = None
    for _ in range(reps):
        model.generate(mel[None], True, tg, ov, True, True)
        t = dict(model.last_timings)
        if best is None or t["ms_loop"] < best["ms_loop"]:
            best = t
    print("%-34s loop %8.1f ms  %6.2f us/step  folds %4d steps %5d" % (
        name, best["ms_loop"], best["ms_loop"] * 1e3 / best["n_steps"], best["n_folds"], best["n_steps"]), flush=True)


cases = [("10s b19 8000/800", 800, 8000, 800), ("60s b137 6000/1000", 4800, 6000, 1000), ("60s b512 1705/170", 4800, 1705, 170),
         ("60s b1024 853/85", 4800, 853, 85)]
for v in ("1", "0"):
    os.environ["WRNN_TC_RAWSAMP"] = v
    for name, T, tg, ov in cases:
        run(raw, "raw rawsamp=%s %s" % (v, name), T, tg, ov)
for name, T, tg, ov in cases[1:]:
    run(mol, "mol %s" % name, T, tg, ov)
