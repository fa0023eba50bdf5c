import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
mol, _ = make_model(seed=12, bits=9, mode="MOL")
mol.precision = 1
mel = norm_mel(400, 1)
try:
    for _ in range(2):
        t0 = time.perf_counter(); mol.generate(mel[None], True, 1705, 170, True, True); dt = time.perf_counter() - t0
    print("OK", os.environ.get("WRNN_TC_COOP"), os.environ.get("WRNN_TC_OVERLAP"), mol.last_timings, dt)
except Exception as ex:
    print("FAIL", os.environ.get("WRNN_TC_COOP"), os.environ.get("WRNN_TC_OVERLAP"), ex)
