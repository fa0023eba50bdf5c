"""numpy restatement of the reference's WaveRNN (fatchord) inference path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference lines it
follows (paths relative to /root/reference).  Arithmetic is float32 where the reference computes in
torch float32 and float64 where it computes in numpy float64.

Pinned against the unmodified reference by tests/test_oracle_golden.py (fixtures minted by
oracle/make_golden.py).  The only deliberate departure is the sampling rule, which BASELINE.json's
north_star asks the build to define: one Philox uniform per (step, fold) and
inverse-CDF for RAW; the same uniforms used exactly as vocoder/distribution.py:123-136 for MOL.
make_golden.py patches the same rule into the reference so both sides are comparable.
"""
import numpy as np

from . import philox
from .weights import (AUX_DIMS, HOP, PAD, RES_BLOCKS, RNN_DIMS, UPSAMPLE)

F32 = np.float32
BN_EPS = F32(1e-5)          # torch.nn.BatchNorm1d default eps
LOG_SCALE_MIN = F32(np.log(1e-14))  # vocoder/distribution.py:113-114


# --------------------------------------------------------------------------------------------------
# conditioning front end
# --------------------------------------------------------------------------------------------------
def pad_mel(mel, pad=PAD):
    """WaveRNN.pad_tensor(side='both') on the time axis, fatchord_version.py:275-288 (called :171).
    mel: (80, T) -> (80, T + 2*pad)."""
    out = np.zeros((mel.shape[0], mel.shape[1] + 2 * pad), F32)
    out[:, pad:pad + mel.shape[1]] = mel
    return out


def _bn(x, sd, p):
    """BatchNorm1d in eval mode (fatchord_version.py:14-15,33; restated wavernn.cpp:294-304)."""
    inv = F32(1.0) / np.sqrt(sd[p + ".running_var"] + BN_EPS)
    return (x - sd[p + ".running_mean"][:, None]) * (inv * sd[p + ".weight"])[:, None] + sd[p + ".bias"][:, None]


def mel_resnet(mel_padded, sd):
    """MelResNet.forward, fatchord_version.py:38-44 with ResBlock :17-24.  (80, T+4) -> (128, T)."""
    w = sd["upsample.resnet.conv_in.weight"]               # (128, 80, 5), no bias, no padding
    k = w.shape[2]
    Tout = mel_padded.shape[1] - k + 1
    x = np.zeros((w.shape[0], Tout), F32)
    for j in range(k):
        x += w[:, :, j] @ mel_padded[:, j:j + Tout]
    x = np.maximum(_bn(x, sd, "upsample.resnet.batch_norm"), 0)
    for i in range(RES_BLOCKS):
        p = "upsample.resnet.layers.%d" % i
        r = x
        x = sd[p + ".conv1.weight"][:, :, 0] @ x
        x = np.maximum(_bn(x, sd, p + ".batch_norm1"), 0)
        x = sd[p + ".conv2.weight"][:, :, 0] @ x
        x = _bn(x, sd, p + ".batch_norm2")
        x = x + r
    x = sd["upsample.resnet.conv_out.weight"][:, :, 0] @ x + sd["upsample.resnet.conv_out.bias"][:, None]
    return x.astype(F32)


def upsample_mel(mel_padded, sd):
    """The three Stretch2d + Conv2d(1,1,(1,2s+1),pad (0,s)) layers and the indent trim,
    fatchord_version.py:47-57,66-75,82-84.  (80, T+4) -> (80, 200*T)."""
    m = mel_padded.astype(F32)
    for idx, s in zip((1, 3, 5), UPSAMPLE):
        m = np.repeat(m, s, axis=1)                              # Stretch2d(x_scale=s, y_scale=1)
        w = sd["upsample.up_layers.%d.weight" % idx].reshape(-1)  # (2s+1,)
        mp = np.pad(m, ((0, 0), (s, s)))
        out = np.zeros_like(m)
        for kk in range(2 * s + 1):                              # cross-correlation, zero padded
            out += w[kk] * mp[:, kk:kk + m.shape[1]]
        m = out
    indent = PAD * HOP
    return m[:, indent:-indent]


def upsample_network(mel, sd):
    """UpsampleNetwork.forward on the padded mel, fatchord_version.py:78-85 (called :171-172).
    mel (80, T) already divided by max_abs_value.  Returns mels (200T, 80), aux (200T, 128)."""
    mp = pad_mel(mel)
    aux = np.repeat(mel_resnet(mp, sd), HOP, axis=1)             # resnet_stretch, :79-81
    m = upsample_mel(mp, sd)
    return np.ascontiguousarray(m.T), np.ascontiguousarray(aux.T)


# --------------------------------------------------------------------------------------------------
# fold / unfold (integer index arithmetic: must be bit-exact)
# --------------------------------------------------------------------------------------------------
def fold_plan(total_len, target, overlap):
    """Index arithmetic of WaveRNN.fold_with_overlap, fatchord_version.py:315-326.
    Returns (num_folds, padded_len)."""
    num_folds = (total_len - overlap) // (target + overlap)
    extended_len = num_folds * (overlap + target) + overlap
    remaining = total_len - extended_len
    padded = total_len
    if remaining != 0:
        num_folds += 1
        padded = total_len + target + 2 * overlap - remaining
    return num_folds, padded


def fold_with_overlap(x, target, overlap):
    """fatchord_version.py:290-340.  x: (total_len, features) -> (num_folds, target+2*overlap, features);
    tail zero padding in the UPSAMPLED domain."""
    total_len, feats = x.shape
    num_folds, padded = fold_plan(total_len, target, overlap)
    xp = np.zeros((padded, feats), x.dtype)
    xp[:total_len] = x
    S = target + 2 * overlap
    folded = np.zeros((num_folds, S, feats), x.dtype)
    for i in range(num_folds):
        start = i * (target + overlap)
        folded[i] = xp[start:start + S]
    return folded


def xfade_and_unfold(y, overlap):
    """fatchord_version.py:342-404 (the `target` argument is ignored there, Q7).  y: (F, S) float64."""
    y = np.array(y, dtype=np.float64)
    num_folds, length = y.shape
    target = length - 2 * overlap
    total_len = num_folds * (target + overlap) + overlap
    silence_len = overlap // 2
    fade_len = overlap - silence_len
    t = np.linspace(-1, 1, fade_len, dtype=np.float64)
    fade_in = np.concatenate([np.zeros(silence_len), np.sqrt(0.5 * (1 + t))])
    fade_out = np.concatenate([np.sqrt(0.5 * (1 - t)), np.zeros(silence_len)])
    if overlap > 0:
        y[:, :overlap] *= fade_in
        y[:, length - overlap:] *= fade_out
    unfolded = np.zeros(total_len, np.float64)
    for i in range(num_folds):
        start = i * (target + overlap)
        unfolded[start:start + length] += y[i]
    return unfolded


# --------------------------------------------------------------------------------------------------
# post chain (host numpy float64 in the reference)
# --------------------------------------------------------------------------------------------------
def decode_mu_law(y, mu):
    """vocoder/audio.py:102-107 with from_labels=False (called fatchord_version.py:247-248)."""
    mu = mu - 1
    return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)


def de_emphasis(x, coef=0.97):
    """vocoder/audio.py:92-93: lfilter([1], [1, -0.97], x), i.e. y[n] = x[n] + 0.97*y[n-1] in float64."""
    y = np.empty_like(x, dtype=np.float64)
    acc = 0.0
    for n in range(x.shape[0]):
        acc = x[n] + coef * acc
        y[n] = acc
    return y


def finish(output, wave_len, n_classes, mu_law, apply_preemphasis):
    """Tail of WaveRNN.generate, fatchord_version.py:247-255."""
    if mu_law:
        output = decode_mu_law(output, n_classes)
    if apply_preemphasis:
        output = de_emphasis(output)
    fade_out = np.linspace(1, 0, 20 * HOP)
    output = output[:wave_len].copy()
    output[-20 * HOP:] *= fade_out      # raises ValueError for T <= 20 exactly like the reference (Q8)
    return output


# --------------------------------------------------------------------------------------------------
# the autoregressive loop
# --------------------------------------------------------------------------------------------------
def _sigmoid(x):
    return F32(1.0) / (F32(1.0) + np.exp(-x))


def gru_cell(x, h, sd, name):
    """torch.nn.GRUCell equations with gate order r,z,n (fatchord_version.py:267-273 builds the cell;
    restated in vocoder/libwavernn/convert.py:207-211 and wavernn.cpp:154-157)."""
    gi = x @ sd[name + ".weight_ih_l0"].T + sd[name + ".bias_ih_l0"]
    gh = h @ sd[name + ".weight_hh_l0"].T + sd[name + ".bias_hh_l0"]
    H = RNN_DIMS
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return ((F32(1.0) - z) * n + z * h).astype(F32)


def step_logits(x, m_t, a_t, h1, h2, sd):
    """One iteration of the loop body, fatchord_version.py:194-213.
    x (B,1), m_t (B,80), a_t (B,128) -> logits (B,C), h1, h2."""
    d = AUX_DIMS
    a1, a2, a3, a4 = (a_t[:, d * i:d * (i + 1)] for i in range(4))
    u = np.concatenate([x, m_t, a1[:, :-1]], axis=1)                 # :198 (drops last aux channel, Q5)
    xI = (u @ sd["I.weight"].T + sd["I.bias"]).astype(F32)          # :199
    h1 = gru_cell(xI, h1, sd, "rnn1")                                # :200
    x1 = xI + h1                                                     # :202
    h2 = gru_cell(np.concatenate([x1, a2], axis=1), h2, sd, "rnn2")  # :203-204
    x2 = x1 + h2                                                     # :206
    f1 = np.maximum(np.concatenate([x2, a3], axis=1) @ sd["fc1.weight"].T + sd["fc1.bias"], 0)  # :207-208
    f2 = np.maximum(np.concatenate([f1, a4], axis=1) @ sd["fc2.weight"].T + sd["fc2.bias"], 0)  # :210-211
    logits = f2 @ sd["fc3.weight"].T + sd["fc3.bias"]                # :213
    return logits.astype(F32), h1, h2


def sample_raw(logits, u):
    """Build-defined RAW rule (north_star; matches libwavernn net_impl.cpp:19-27,129-140): softmax in
    float32 (fatchord_version.py:225), sequential float32 cumulative sum, first k with cdf[k] >= u;
    clamped to C-1 when rounding leaves cdf[C-1] < u.  Returns int64 class indices (B,)."""
    l = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(l).astype(F32)
    p = (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    cdf = np.cumsum(p, axis=1, dtype=F32)
    k = (cdf < u[:, None]).sum(axis=1)
    return np.minimum(k, logits.shape[1] - 1).astype(np.int64)


def sample_mol(logits, u_mix, u_log):
    """vocoder/distribution.py:104-140 with the two uniform_() draws replaced by injected noise.
    logits (B,30); u_mix (B,10), u_log (B,) already in [1e-5, 1-1e-5].  Returns (x (B,), k (B,))."""
    nr = logits.shape[1] // 3
    temp = logits[:, :nr] - np.log(-np.log(u_mix))                    # :123-124
    k = temp.argmax(axis=1)                                           # :125
    rows = np.arange(logits.shape[0])
    means = logits[rows, nr + k]                                      # :130
    log_scales = np.maximum(logits[rows, 2 * nr + k], LOG_SCALE_MIN)  # :131-132
    x = means + np.exp(log_scales) * (np.log(u_log) - np.log(F32(1.0) - u_log))   # :135-136
    return np.clip(x, -1.0, 1.0).astype(F32), k                      # :138


def label_to_float(k, C):
    """fatchord_version.py:228 in float32 (Q10): 2*k/(C-1) - 1."""
    return (F32(2.0) * k.astype(F32) / F32(C - 1.0) - F32(1.0)).astype(F32)


def generate(mel, sd, mode="RAW", batched=True, target=8000, overlap=800, mu_law=True,
             apply_preemphasis=True, seed=0, utt=0, forced_samples=None, return_trace=False,
             max_steps=None):
    """WaveRNN.generate, fatchord_version.py:155-259.  mel (80, T) float32 already normalised.

    forced_samples: optional (F, S) float32 array fed back instead of the oracle's own samples
    (index-teacher-forcing, SURVEY.md section 7 "Sampling parity").
    Returns wav float64[(T-1)*200]; with return_trace also a dict (logits, samples, indices/mixture).
    """
    C = sd["fc3.weight"].shape[0]
    mu_law = mu_law if mode == "RAW" else False                      # :156
    T = mel.shape[1]
    wave_len = (T - 1) * HOP                                         # :170
    mels, aux = upsample_network(mel.astype(F32), sd)                # :171-172
    if batched:
        mels = fold_with_overlap(mels, target, overlap)              # :174-176
        aux = fold_with_overlap(aux, target, overlap)
    else:
        mels, aux = mels[None], aux[None]
    B, S, _ = mels.shape
    if max_steps is not None:
        S = min(S, max_steps)
    h1 = np.zeros((B, RNN_DIMS), F32)
    h2 = np.zeros((B, RNN_DIMS), F32)
    x = np.zeros((B, 1), F32)                                        # :178-187
    if mode == "RAW":
        U = philox.raw_uniforms(seed, S, B, utt=utt)
    else:
        UM, UL = philox.mol_uniforms(seed, S, B, utt=utt)
    out = np.zeros((B, S), F32)
    tr_logits = np.zeros((B, S, C), F32) if return_trace else None
    tr_idx = np.zeros((B, S), np.int64)
    for i in range(S):                                               # :192
        logits, h1, h2 = step_logits(x, mels[:, i], aux[:, i], h1, h2, sd)
        if return_trace:
            tr_logits[:, i] = logits
        if mode == "MOL":
            s, k = sample_mol(logits, UM[i], UL[i])                  # :215-222
        elif mode == "RAW":
            k = sample_raw(logits, U[i])                             # :224-230
            s = label_to_float(k, C)
        else:
            raise RuntimeError("Unknown model mode value - ", mode)  # :232
        out[:, i] = s
        tr_idx[:, i] = k
        x = (forced_samples[:, i] if forced_samples is not None else s).reshape(B, 1).astype(F32)
    trace = {"logits": tr_logits, "samples": out.copy(), "index": tr_idx}
    if max_steps is not None:
        return None, trace
    output = out.astype(np.float64)                                  # :238-240
    output = xfade_and_unfold(output, overlap) if batched else output[0]   # :242-245
    wav = finish(output, wave_len, C, mu_law, apply_preemphasis)     # :247-255
    return (wav, trace) if return_trace else wav


def teacher_forced_logits(x_seq, mel, sd):
    """WaveRNN.forward in eval mode, fatchord_version.py:120-153 (row a17): x_seq (200T,) float32 is
    the previous-sample input at every step; returns logits (200T, C)."""
    mels, aux = upsample_network(mel.astype(F32), sd)
    N = mels.shape[0]
    C = sd["fc3.weight"].shape[0]
    h1 = np.zeros((1, RNN_DIMS), F32)
    h2 = np.zeros((1, RNN_DIMS), F32)
    out = np.zeros((N, C), F32)
    for i in range(N):
        x = np.array([[x_seq[i]]], F32)
        out[i], h1, h2 = step_logits(x, mels[i:i + 1], aux[i:i + 1], h1, h2, sd)
    return out


# --------------------------------------------------------------------------------------------------
# libwavernn wire format (vocoder/libwavernn/convert.py:61-84)
# --------------------------------------------------------------------------------------------------
def compress(W, group=4):
    """convert.compress: kept 1 x group blocks (row-major) and the uint8 group-column index stream with
    255 as row-end marker (rows+1 markers, convert.py:70-73)."""
    N = W.shape[1]
    nz = (W != 0).reshape(W.shape[0], N // group, group).max(axis=-1)
    row, col = np.nonzero(nz)
    idx = []
    for i in range(nz.shape[0] + 1):
        idx += list(col[row == i])
        idx += [255]
    mask = np.repeat(nz, group, axis=1)
    return W[mask].astype(F32), np.asarray(idx, dtype=np.uint8)
