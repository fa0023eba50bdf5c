"""CPU restatement (numpy) of the `runtimeracer-wavernn` topology -- SURVEY.md section 8(f) row 1, the variant libwavernn's
build.sh builds and `load_model` hard-codes for the C++ path (vocoder/inference.py:43).  TEST INFRASTRUCTURE ONLY: groundwork
for the next hot-path row; nothing in the product uses this topology yet.

Reference: vocoder/models/runtimeracer_version.py -- constructor :97-134 (rnn_dims = fc_dims = 256, aux_dims =
res_out_dims // 4 = 32, four GRUs, five FC layers), generate() body :248-268, sampling :270-288 (identical to the fatchord
rules), conditioning / fold / crossfade / post chain shared with the fatchord model (same classes, same hparams).
Pinned against tests/golden/rr_*.npz (minted from the unmodified reference by oracle/make_golden_rr.py).
"""
import numpy as np

from . import philox
from . import wavernn_oracle as orc
from .weights import AUX_DIMS, FEAT_DIMS, HOP, make_state_dict, n_classes

F32 = np.float32
RR_RNN = 256      # config/hparams.py:363
RR_FC = 256       # :364


from .weights import make_state_dict_rr  # noqa: E402,F401  (the generator lives with the other input generators: rtvc_b200/synth.py)


def _gru(x, h, sd, name):
    """torch.nn.GRUCell, gate order r, z, n (runtimeracer_version.py get_gru_cell; same equations as orc.gru_cell, H = 256)."""
    gi = x @ sd[name + ".weight_ih_l0"].T + sd[name + ".bias_ih_l0"]
    gh = h @ sd[name + ".weight_hh_l0"].T + sd[name + ".bias_hh_l0"]
    H = RR_RNN
    sig = lambda v: F32(1.0) / (F32(1.0) + np.exp(-v))
    r = sig(gi[:, :H] + gh[:, :H])
    z = sig(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return ((F32(1.0) - z) * n + z * h).astype(F32)


def step_logits_rr(x, m_t, a_t, hs, sd):
    """One iteration of runtimeracer_version.py:248-268.  x (B,1), m_t (B,80), a_t (B,128), hs = [h1..h4] -> logits, hs."""
    d = AUX_DIMS
    a1, a2, a3, a4 = (a_t[:, d * i:d * (i + 1)] for i in range(4))
    h1, h2, h3, h4 = hs
    v = np.concatenate([x, m_t, a1[:, :-1]], axis=1)                               # :250
    v = (v @ sd["I.weight"].T + sd["I.bias"]).astype(F32)                          # :251
    h1 = _gru(v, h1, sd, "rnn1"); v = v + h1                                       # :253-254
    h2 = _gru(v, h2, sd, "rnn2"); v = v + h2                                       # :255-256
    h3 = _gru(np.concatenate([v, a2], axis=1), h3, sd, "rnn3"); v = v + h3         # :258-260
    h4 = _gru(v, h4, sd, "rnn4"); v = v + h4                                       # :261-262
    v = np.concatenate([v, a3], axis=1) @ sd["fc1.weight"].T + sd["fc1.bias"]      # :264-265 (no activation after fc1)
    v = np.maximum(v @ sd["fc2.weight"].T + sd["fc2.bias"], 0)                     # :266
    v = np.concatenate([v.astype(F32), a4], axis=1) @ sd["fc3.weight"].T + sd["fc3.bias"]   # :268-269 (no activation after fc3)
    v = np.maximum(v @ sd["fc4.weight"].T + sd["fc4.bias"], 0)                     # :270
    logits = v @ sd["fc5.weight"].T + sd["fc5.bias"]                               # :272
    return logits.astype(F32), [h1, h2, h3, h4]


def generate_rr(mel_norm, sd, seed, mode="RAW", bits=9, batched=True, target=8000, overlap=800, mu_law=True, preemph=True,
                forced=None, max_steps=0):
    """runtimeracer_version.py:generate with the build's Philox noise contract (oracle/philox.py).  Returns dict(wav, logits
    (B,S,C), samples (B,S))."""
    C = n_classes(bits, mode)
    T = mel_norm.shape[1]
    mels, aux = orc.upsample_network(mel_norm, sd)
    if batched:
        mels, aux = orc.fold_with_overlap(mels, target, overlap), orc.fold_with_overlap(aux, target, overlap)
    else:
        mels, aux = mels[None], aux[None]
    B, S, _ = mels.shape
    if max_steps:
        S = min(S, max_steps)
    hs = [np.zeros((B, RR_RNN), F32) for _ in range(4)]
    x = np.zeros((B, 1), F32)
    logits = np.zeros((B, S, C), F32)
    samples = np.zeros((B, S), F32)
    if mode == "RAW":
        U = philox.raw_uniforms(seed, S, B)
    else:
        UM, UL = philox.mol_uniforms(seed, S, B)
    for i in range(S):
        lg, hs = step_logits_rr(x, mels[:, i], aux[:, i], hs, sd)
        logits[:, i] = lg
        if mode == "RAW":
            samples[:, i] = orc.label_to_float(orc.sample_raw(lg, U[i]), C)
        else:
            samples[:, i] = orc.sample_mol(lg, UM[i], UL[i])[0]
        x = (forced[:, i] if forced is not None else samples[:, i]).reshape(B, 1).astype(F32)
    out = dict(logits=logits, samples=samples, wav=None)
    if not max_steps:
        y = samples.astype(np.float64)
        y = orc.xfade_and_unfold(y, overlap) if batched else y[0]
        out["wav"] = orc.finish(y, (T - 1) * HOP, C, mu_law and mode == "RAW", preemph)
    return out
