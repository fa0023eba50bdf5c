"""CPU restatement (numpy) of the `geneing-wavernn` topology -- SURVEY.md section 8(f) row 3 (one GRU-256, two FC layers, aux split in
two; hop 200 = 4 x 5 x 10, MelResNet with 64 channels and 3 blocks).  TEST INFRASTRUCTURE ONLY: groundwork, no product path yet.

Reference: vocoder/models/geneing_version.py -- constructor :89-121, generate() body :186-232 (mode 'BITS' = softmax over 2**bits
classes, sampled like the fatchord RAW mode; 'MOL' as fatchord; the beta-distribution mode 'RAW' is not restated), hparams
config/hparams.py:288-300.  The front end is the same UpsampleNetwork class with other sizes, so it is restated here generically
(sizes read from the state_dict).  Pinned against tests/golden/gn_bits9.npz (oracle/make_golden_gn.py).
"""
import numpy as np

from . import philox
from . import wavernn_oracle as orc

F32 = np.float32
GN = dict(rnn_dims=256, fc_dims=128, compute_dims=64, res_out_dims=64, res_blocks=3, upsample=(4, 5, 10), pad=2, feat=80)
HOP = 200


from .weights import make_state_dict_gn  # noqa: E402,F401  (rtvc_b200/synth.py)


def upsample_network_generic(mel, sd, pad=2):
    """UpsampleNetwork.forward (geneing_version.py, same class as fatchord_version.py:60-85) for any sizes: (80, T) ->
    mels (hop*T, 80), aux (hop*T, res_out_dims)."""
    n_blocks = sum(1 for k in sd if k.startswith("upsample.resnet.layers.") and k.endswith(".conv1.weight"))
    factors = [(sd["upsample.up_layers.%d.weight" % i].size - 1) // 2 for i in (1, 3, 5)]
    hop = int(np.prod(factors))
    mp = np.pad(mel.astype(F32), ((0, 0), (pad, pad)))
    w = sd["upsample.resnet.conv_in.weight"]
    Tout = mp.shape[1] - w.shape[2] + 1
    x = np.zeros((w.shape[0], Tout), F32)
    for j in range(w.shape[2]):
        x += w[:, :, j] @ mp[:, j:j + Tout]
    x = np.maximum(orc._bn(x, sd, "upsample.resnet.batch_norm"), 0)
    for i in range(n_blocks):
        p = "upsample.resnet.layers.%d" % i
        r = x
        x = np.maximum(orc._bn(sd[p + ".conv1.weight"][:, :, 0] @ x, sd, p + ".batch_norm1"), 0)
        x = orc._bn(sd[p + ".conv2.weight"][:, :, 0] @ x, sd, p + ".batch_norm2") + r
    x = (sd["upsample.resnet.conv_out.weight"][:, :, 0] @ x + sd["upsample.resnet.conv_out.bias"][:, None]).astype(F32)
    aux = np.repeat(x, hop, axis=1)
    m = mp
    for idx, s in zip((1, 3, 5), factors):
        m = np.repeat(m, s, axis=1)
        wk = sd["upsample.up_layers.%d.weight" % idx].reshape(-1)
        mpad = np.pad(m, ((0, 0), (s, s)))
        out = np.zeros_like(m)
        for kk in range(2 * s + 1):
            out += wk[kk] * mpad[:, kk:kk + m.shape[1]]
        m = out
    indent = pad * hop
    return np.ascontiguousarray(m[:, indent:-indent].T), np.ascontiguousarray(aux.T)


def step_logits_gn(x, m_t, a_t, h1, sd):
    """One iteration of geneing_version.py:199-210.  x (B,1), m_t (B,80), a_t (B,64) -> logits (B,C), h1."""
    d = a_t.shape[1] // 2
    a1, a2 = a_t[:, :d], a_t[:, d:2 * d]
    v = (np.concatenate([x, m_t, a1[:, :-1]], axis=1) @ sd["I.weight"].T + sd["I.bias"]).astype(F32)      # :203-204
    gi = v @ sd["rnn1.weight_ih_l0"].T + sd["rnn1.bias_ih_l0"]
    gh = h1 @ sd["rnn1.weight_hh_l0"].T + sd["rnn1.bias_hh_l0"]
    H = h1.shape[1]
    sig = lambda q: F32(1.0) / (F32(1.0) + np.exp(-q))
    r, z = sig(gi[:, :H] + gh[:, :H]), sig(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    h1 = ((F32(1.0) - z) * n + z * h1).astype(F32)                                                             # :205
    v = np.concatenate([v + h1, a2], axis=1)                                                                  # :207-208
    v = np.maximum(v @ sd["fc1.weight"].T + sd["fc1.bias"], 0)                                               # :209
    return (v @ sd["fc3.weight"].T + sd["fc3.bias"]).astype(F32), h1                                         # :211


def generate_gn(mel_norm, sd, seed, bits=9, batched=True, target=1000, overlap=200, forced=None, max_steps=0, mode="BITS"):
    """geneing_version.py:generate in mode 'BITS' (mu_law False, config/hparams.py:292) or 'MOL' (:217-223, the fatchord rule) with the
    build's Philox noise contract."""
    C = 2 ** bits if mode == "BITS" else 30
    T = mel_norm.shape[1]
    mels, aux = upsample_network_generic(mel_norm, sd)
    if batched:
        mels, aux = orc.fold_with_overlap(mels, target, overlap), orc.fold_with_overlap(aux, target, overlap)
    else:
        mels, aux = mels[None], aux[None]
    B, S, _ = mels.shape
    if max_steps:
        S = min(S, max_steps)
    h1 = np.zeros((B, GN["rnn_dims"]), F32)
    x = np.zeros((B, 1), F32)
    logits, samples = np.zeros((B, S, C), F32), np.zeros((B, S), F32)
    if mode == "BITS":
        U = philox.raw_uniforms(seed, S, B)
    else:
        UM, UL = philox.mol_uniforms(seed, S, B)
    for i in range(S):
        lg, h1 = step_logits_gn(x, mels[:, i], aux[:, i], h1, sd)
        logits[:, i] = lg
        if mode == "BITS":
            samples[:, i] = orc.label_to_float(orc.sample_raw(lg, U[i]), C)
        else:
            samples[:, i] = orc.sample_mol(lg, UM[i], UL[i])[0]
        x = (forced[:, i] if forced is not None else samples[:, i]).reshape(B, 1).astype(F32)
    out = dict(logits=logits, samples=samples, wav=None)
    if not max_steps:
        y = samples.astype(np.float64)
        y = orc.xfade_and_unfold(y, overlap) if batched else y[0]
        out["wav"] = orc.finish(y, (T - 1) * HOP, C, False, True)
    return out
