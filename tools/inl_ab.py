"""A/B of the role-specialised loop's conditioning: inline (W_q . m inside the MMA, per-frame rows) vs records from expander CTAs.
   python tools/inl_ab.py [parity] [time]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, golden, norm_mel

def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
what = sys.argv[1:] or ["parity", "time"]
mol, _ = make_model(seed=12, bits=9, mode="MOL")
raw, _ = make_model(seed=11, bits=9, mode="RAW")

def setenv(**kw):
    for k, v in kw.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)

if "parity" in what:
    g = golden("gen_mol_batched.npz")
    mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
    ref = g["samples"]
    B, Sm1 = ref.shape
    forced = np.zeros((B, Sm1 + 1), np.float32); forced[:, :-1] = ref
    for inl in (0, 1):
        setenv(WRNN_RS_INLINE=inl)
        t0 = time.time()
        out = mol.generate_debug(mel, True, int(g["target"]), int(g["overlap"]), forced=forced, want_logits=True, seed=int(g["seed"]), precision=1)
        lg = out["logits"][:, ::8]
        m = min(lg.shape[1], g["logits_sub"].shape[1])
        k = min(out["logits"].shape[1], Sm1)
        d = np.abs(out["samples"][:, :k] - ref[:, :k])
        print("MOL inl=%d: %s %d folds x %d steps %.2f s: logits rel err vs golden %.3e, samples within 1e-3 %.5f" % (
            inl, dict(mol.last_timings).get("loop_kernel"), B, k, time.time() - t0, rel(lg[:, :m], g["logits_sub"][:, :m]), float((d < 1e-3).mean())), flush=True)
    # RAW, free running, against the fp32 loop; 37 and 213 folds
    for T, tg, ov, steps in ((92, 400, 100, 300), (4800, 3000, 1500, 400)):
        melr = norm_mel(T, 5)
        f32 = raw.generate_debug(melr, True, tg, ov, want_logits=True, seed=3, max_steps=steps, precision=0)
        fo = np.pad(f32["samples"], ((0, 0), (0, tg + 2 * ov - f32["samples"].shape[1])))
        for inl in (0, 1):
            setenv(WRNN_RS_INLINE=inl)
            out = raw.generate_debug(melr, True, tg, ov, forced=fo, want_logits=True, seed=3, max_steps=steps, precision=1)
            same = float((out["samples"] == f32["samples"]).mean())
            print("RAW inl=%d: %s %d folds x %d steps: logits rel err vs fp32 loop %.3e, identical draws %.5f" % (
                inl, dict(raw.last_timings).get("loop_kernel"), out["samples"].shape[0], steps, rel(out["logits"], f32["logits"]), same), flush=True)

if "time" in what:
    def run(model, name, T, tg, ov, reps=2):
        mel = norm_mel(T, 1)
        model.precision = 1
        best = None
        for _ in range(reps):
            model.generate(mel[None], True, tg, ov, True, True)
            t = dict(model.last_timings)
            if best is None or t["ms_loop"] < best["ms_loop"]: best = t
        print("%-34s %s loop %.1f ms  %.2f us/step  folds %d steps %d" % (name, best.get("loop_kernel"), best["ms_loop"], best["ms_loop"] * 1e3 / best["n_steps"],
                                                                        best["n_folds"], best["n_steps"]), flush=True)
    for inl, groups in ((0, None), (1, 2), (1, 3), (1, None)):
        setenv(WRNN_RS_INLINE=inl, WRNN_RS_GROUPS=groups)
        tag = "inl=%d groups=%s " % (inl, groups)
        run(mol, tag + "cfg3ref mol b213", 4800, 3000, 1500)
        run(mol, tag + "cfg3a mol b137", 4800, 6000, 1000)
        run(raw, tag + "cfg1 raw9 b19", 800, 8000, 800)
        run(raw, tag + "raw9 60s b137", 4800, 6000, 1000)
