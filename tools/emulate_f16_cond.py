"""CPU emulation of the fp16 roundings of the role-specialised loop, teacher-forced on the oracle's own samples:
  A: weights + travelling activations in fp16 (what loop_rs.cu does), conditioning in fp32
  B: A + the MEL share of the conditioning as an fp16 x fp16 product (W_q . m_t inside the MMA, K = 80)
  C: B with m_t split in hi + lo fp16 (K = 160)
Prints the logits error (max |d| / max |ref|) and the fraction of identical draws of each against the fp32 oracle.
   python tools/emulate_f16_cond.py [MOL|RAW] [folds] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wavernn_oracle as O, weights, philox

F32 = np.float32
mode = sys.argv[1] if len(sys.argv) > 1 else "MOL"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 600
sd = weights.make_state_dict(seed=12, bits=9, mode=mode)
mel = weights.synthetic_mel(B * 4 + 12, seed=3) / F32(4.0)
mels, aux = O.upsample_network(mel.astype(F32), sd)
N = mels.shape[0]
st = (N - S) // B
mels = np.stack([mels[i * st:i * st + S] for i in range(B)])
aux = np.stack([aux[i * st:i * st + S] for i in range(B)])
C = sd["fc3.weight"].shape[0]
H = 512

def q(x): return np.asarray(x, F32).astype(np.float16).astype(F32)

# reference run (fp32 oracle), free-running
if mode == "RAW":
    U = philox.raw_uniforms(5, S, B, utt=0)
else:
    UM, UL = philox.mol_uniforms(5, S, B, utt=0)
def draw(lg, i):
    if mode == "RAW":
        k = O.sample_raw(lg, U[i]); return O.label_to_float(k, C), k
    s, k = O.sample_mol(lg, UM[i], UL[i]); return s, s
h1 = np.zeros((B, H), F32); h2 = np.zeros((B, H), F32); x = np.zeros((B, 1), F32)
ref_l = np.zeros((S, B, C), F32); ref_x = np.zeros((S, B), F32); ref_k = []
for i in range(S):
    lg, h1, h2 = O.step_logits(x, mels[:, i], aux[:, i], h1, h2, sd)
    s, k = draw(lg, i)
    ref_l[i] = lg; ref_x[i] = s; ref_k.append(k)
    x = s.reshape(B, 1).astype(F32)

I = sd["I.weight"].astype(np.float64); bI = sd["I.bias"].astype(np.float64)
Wi1 = sd["rnn1.weight_ih_l0"].astype(np.float64); Wh1 = sd["rnn1.weight_hh_l0"]
Wi2 = sd["rnn2.weight_ih_l0"].astype(np.float64); Wh2 = sd["rnn2.weight_hh_l0"]
F1 = sd["fc1.weight"].astype(np.float64); F2 = sd["fc2.weight"]; F3 = sd["fc3.weight"]
Wi2a, Wi2b = Wi2[:, :H], Wi2[:, H:]
F1a, F1b = F1[:, :H], F1[:, H:]
F2a, F2b = F2[:, :H], F2[:, H:]
# folded input-side matrices (engine.cu wrnn_finalize): [x | mel 80 | a1 31 | 1]
IB = np.concatenate([I, bI[:, None]], axis=1)
P1, P2, P3 = Wi1 @ IB, Wi2a @ IB, F1a @ IB
sig = lambda v: 1.0 / (1.0 + np.exp(-v))

def run(variant):
    h1 = np.zeros((B, H), F32); h2 = np.zeros((B, H), F32)
    out = np.zeros((S, B, C), F32); same = 0; tot = 0
    for i in range(S):
        x = (ref_x[i - 1] if i > 0 else np.zeros(B, F32)).reshape(B, 1).astype(np.float64)
        m = mels[:, i].astype(np.float64); a = aux[:, i].astype(np.float64)
        a1, a2, a3, a4 = a[:, :31], a[:, 32:64], a[:, 64:96], a[:, 96:128]
        def mel_term(P):
            Wq = P[:, 1:81]
            if variant == "A": return m @ Wq.T
            if variant == "B": return q(m).astype(np.float64) @ q(Wq).astype(np.float64).T
            hi = q(m); lo = q(m - hi)
            return (hi.astype(np.float64) + lo) @ q(Wq).astype(np.float64).T
        def rest(P): return x * P[:, 0] + a1 @ P[:, 81:112].T + P[:, 112]
        gi1 = rest(P1) + mel_term(P1) + sd["rnn1.bias_ih_l0"]
        gh1 = q(h1).astype(np.float64) @ q(Wh1).astype(np.float64).T + sd["rnn1.bias_hh_l0"]
        r = sig(gi1[:, :H] + gh1[:, :H]); z = sig(gi1[:, H:2 * H] + gh1[:, H:2 * H]); n = np.tanh(gi1[:, 2 * H:] + r * gh1[:, 2 * H:])
        h1 = ((1 - z) * n + z * h1).astype(F32)
        gi2 = rest(P2) + mel_term(P2) + q(h1).astype(np.float64) @ q(Wi2a).astype(np.float64).T + a2 @ Wi2b.T + sd["rnn2.bias_ih_l0"]
        gh2 = q(h2).astype(np.float64) @ q(Wh2).astype(np.float64).T + sd["rnn2.bias_hh_l0"]
        r = sig(gi2[:, :H] + gh2[:, :H]); z = sig(gi2[:, H:2 * H] + gh2[:, H:2 * H]); n = np.tanh(gi2[:, 2 * H:] + r * gh2[:, 2 * H:])
        h2 = ((1 - z) * n + z * h2).astype(F32)
        s2 = q(np.clip(q(h1) + h2, -1.999, 1.999)).astype(np.float64)
        f1 = np.maximum(rest(P3) + mel_term(P3) + s2 @ q(F1a).astype(np.float64).T + a3 @ F1b.T + sd["fc1.bias"], 0).astype(F32)
        f2 = np.maximum(q(f1).astype(np.float64) @ q(F2a).astype(np.float64).T + a4 @ F2b.astype(np.float64).T + sd["fc2.bias"], 0).astype(F32)
        lg = (q(f2).astype(np.float64) @ q(F3).astype(np.float64).T + sd["fc3.bias"]).astype(F32)
        out[i] = lg
        s, k = draw(lg, i)
        if mode == "RAW": same += int((k == ref_k[i]).sum())
        else: same += int((np.abs(s - ref_k[i]) < 1e-3).sum())
        tot += B
    return float(np.abs(out - ref_l).max() / np.abs(ref_l).max()), same / tot

for v in ("A", "B", "C"):
    e, ag = run(v)
    print("%s  variant %s: logits rel err %.3e  draws identical %.5f  (%d folds x %d steps)" % (mode, v, e, ag, B, S), flush=True)
