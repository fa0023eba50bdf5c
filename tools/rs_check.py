"""Quick parity / timing check of the role-specialised loop (loop_rs.cu) against the golden MOL run and loop_tc.cu."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, golden, norm_mel

F16 = 1
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())

model, _ = make_model(seed=12, bits=9, mode="MOL")
g = golden("gen_mol_batched.npz")
mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
ref = g["samples"]
B, Sm1 = ref.shape
forced = np.zeros((B, Sm1 + 1), np.float32)
forced[:, :-1] = ref
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 0
t0 = time.time()
out = model.generate_debug(mel, True, int(g["target"]), int(g["overlap"]), forced=forced, want_logits=True, seed=int(g["seed"]), precision=F16,
                           max_steps=steps)
print("rs: %d folds x %d steps in %.2f s" % (out["samples"].shape[0], out["samples"].shape[1], time.time() - t0), flush=True)
n = out["logits"].shape[1]
lg = out["logits"][:, ::8]
m = min(lg.shape[1], g["logits_sub"].shape[1])
print("logits rel err vs reference golden: %.3e" % rel(lg[:, :m], g["logits_sub"][:, :m]))
k = min(n, Sm1)
d = np.abs(out["samples"][:, :k] - ref[:, :k])
print("samples within 1e-3: %.5f (max %.3e)" % (float((d < 1e-3).mean()), float(d.max())))
os.environ["WRNN_RS"] = "0"
out2 = model.generate_debug(mel, True, int(g["target"]), int(g["overlap"]), forced=forced, want_logits=True, seed=int(g["seed"]), precision=F16,
                            max_steps=steps)
print("vs loop_tc: logits rel %.3e, samples max diff %.3e" % (rel(out["logits"], out2["logits"]), float(np.abs(out["samples"] - out2["samples"]).max())))
