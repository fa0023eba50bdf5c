"""Step time of the tensor-core loop vs. (fold count, fold sets per group): separates per-tile costs from bandwidth."""
import sys, os, time, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from tests.util import make_model, norm_mel
    tg, ov = int(sys.argv[1]), int(sys.argv[2])
    mol, _ = make_model(seed=12, bits=9, mode=sys.argv[3] if len(sys.argv) > 3 else "MOL")
    mol.precision = 1
    mel = norm_mel(4800, 1)
    for _ in range(2):
        mol.generate(mel[None], True, tg, ov, True, True)
    t = mol.last_timings
    print("sets=%s tg=%d ov=%d folds %d steps %d  loop %.1f ms  %.2f us/step  %.2f folds/us" % (
        os.environ.get("WRNN_TC_SETS", "auto"), tg, ov, t["n_folds"], t["n_steps"], t["ms_loop"], t["ms_loop"] * 1e3 / t["n_steps"],
        t["n_folds"] / (t["ms_loop"] * 1e3 / t["n_steps"])), flush=True)
else:
    for cl in ("1", "4"):
        for fl in ("0", "1"):
            for tg, ov, sets in [(3410, 341, 1), (1705, 170, 2), (853, 85, 4)]:
                env = dict(os.environ, WRNN_TC_SETS=str(sets), WRNN_TC_CLUSTER=cl, WRNN_TC_FLAGS=fl)
                print("cluster", cl, "inline-release", fl, end="  ", flush=True)
                subprocess.run([sys.executable, __file__, str(tg), str(ov)], env=env)
