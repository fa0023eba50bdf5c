"""Deterministic random-init fatchord WaveRNN weights and synthetic mels (numpy only, reproducible from a seed).

Input generation for benchmarks, smoke tests and the test-suite (no arithmetic of the path lives here).  Produces a state_dict with exactly the keys/shapes the reference model
has (vocoder/models/fatchord_version.py:88-118; layout written by vocoder/train.py:316-324), so the
same weights can be loaded into the reference (oracle/make_golden.py), the numpy oracle and the CUDA
engine without shipping 18 MB fixtures.  Scales follow torch's default inits (Linear/GRU:
U(+-1/sqrt(fan)), conv: U(+-1/sqrt(fan_in))); BatchNorm statistics are randomised and the upsample
filters perturbed so that nothing on the path is an identity (SURVEY.md section 8d "Weights").
"""
import numpy as np

RNN_DIMS = 512
FC_DIMS = 512
FEAT_DIMS = 80
COMPUTE_DIMS = 128
RES_OUT_DIMS = 128
RES_BLOCKS = 10
AUX_DIMS = RES_OUT_DIMS // 4
UPSAMPLE = (5, 5, 8)
PAD = 2
HOP = 200


def n_classes(bits, mode):
    return 2 ** bits if mode == "RAW" else 30


def make_state_dict(seed=0, bits=9, mode="RAW", logit_gain=1.0):
    """dict name -> np.ndarray (float32, int64 for step / num_batches_tracked)."""
    rng = np.random.default_rng(seed)

    def U(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    sd = {"step": np.zeros((1,), np.int64)}

    def bn(prefix):
        sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, COMPUTE_DIMS).astype(np.float32)
        sd[prefix + ".bias"] = (0.2 * rng.standard_normal(COMPUTE_DIMS)).astype(np.float32)
        sd[prefix + ".running_mean"] = (0.3 * rng.standard_normal(COMPUTE_DIMS)).astype(np.float32)
        sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, COMPUTE_DIMS).astype(np.float32)
        sd[prefix + ".num_batches_tracked"] = np.zeros((), np.int64)

    k = 2 * PAD + 1
    sd["upsample.resnet.conv_in.weight"] = U((COMPUTE_DIMS, FEAT_DIMS, k), FEAT_DIMS * k)
    bn("upsample.resnet.batch_norm")
    for i in range(RES_BLOCKS):
        p = "upsample.resnet.layers.%d" % i
        sd[p + ".conv1.weight"] = U((COMPUTE_DIMS, COMPUTE_DIMS, 1), COMPUTE_DIMS)
        sd[p + ".conv2.weight"] = U((COMPUTE_DIMS, COMPUTE_DIMS, 1), COMPUTE_DIMS)
        bn(p + ".batch_norm1")
        bn(p + ".batch_norm2")
    sd["upsample.resnet.conv_out.weight"] = U((RES_OUT_DIMS, COMPUTE_DIMS, 1), COMPUTE_DIMS)
    sd["upsample.resnet.conv_out.bias"] = U((RES_OUT_DIMS,), COMPUTE_DIMS)
    for idx, s in zip((1, 3, 5), UPSAMPLE):
        w = np.full((1, 1, 1, 2 * s + 1), 1.0 / (2 * s + 1), np.float64)
        w = w + 0.02 * rng.standard_normal(w.shape)
        sd["upsample.up_layers.%d.weight" % idx] = w.astype(np.float32)

    C = n_classes(bits, mode)
    n_in = FEAT_DIMS + AUX_DIMS - 1 + 1
    sd["I.weight"] = U((RNN_DIMS, n_in), n_in)
    sd["I.bias"] = U((RNN_DIMS,), n_in)
    for name, n_inp in (("rnn1", RNN_DIMS), ("rnn2", RNN_DIMS + AUX_DIMS)):
        sd[name + ".weight_ih_l0"] = U((3 * RNN_DIMS, n_inp), RNN_DIMS)
        sd[name + ".weight_hh_l0"] = U((3 * RNN_DIMS, RNN_DIMS), RNN_DIMS)
        sd[name + ".bias_ih_l0"] = U((3 * RNN_DIMS,), RNN_DIMS)
        sd[name + ".bias_hh_l0"] = U((3 * RNN_DIMS,), RNN_DIMS)
    sd["fc1.weight"] = U((FC_DIMS, RNN_DIMS + AUX_DIMS), RNN_DIMS + AUX_DIMS)
    sd["fc1.bias"] = U((FC_DIMS,), RNN_DIMS + AUX_DIMS)
    sd["fc2.weight"] = U((FC_DIMS, FC_DIMS + AUX_DIMS), FC_DIMS + AUX_DIMS)
    sd["fc2.bias"] = U((FC_DIMS,), FC_DIMS + AUX_DIMS)
    sd["fc3.weight"] = (logit_gain * U((C, FC_DIMS), FC_DIMS)).astype(np.float32)
    sd["fc3.bias"] = (logit_gain * U((C,), FC_DIMS)).astype(np.float32)
    return sd


def prune_state_dict(sd, z=0.9, group=4):
    """Apply the reference's magnitude pruning once (vocoder/pruner.py:60-88, layers per
    fatchord_version.py:115, GRU ih+hh per pruner.py:29-30): per gate block, zero the k smallest
    1 x group column groups by sum(|w|); ties at the threshold are kept (S >= threshold)."""
    out = dict(sd)

    def mask_block(W):
        rows, cols = W.shape
        S = np.abs(W).reshape(rows, cols // group, group).sum(axis=2)
        flat = np.sort(S.reshape(-1), kind="stable")
        kk = int(rows * cols // group * z)
        thr = flat[kk]
        M = (S >= thr).astype(np.float32)
        return W * np.repeat(M, group, axis=1)

    for name, splits in (("I.weight", 1), ("rnn1.weight_ih_l0", 3), ("rnn1.weight_hh_l0", 3),
                         ("rnn2.weight_ih_l0", 3), ("rnn2.weight_hh_l0", 3),
                         ("fc1.weight", 1), ("fc2.weight", 1), ("fc3.weight", 1)):
        W = sd[name]
        blocks = np.split(W, splits, axis=0)
        out[name] = np.concatenate([mask_block(b) for b in blocks], axis=0).astype(np.float32)
    return out


def synthetic_mel(T, seed=1, lo=-4.0, hi=4.0):
    """(80, T) float32 mel in the synthesizer's range (synthesizer/inference.py:96-97)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, size=(FEAT_DIMS, T)).astype(np.float32)


# ---- the other two topologies of the reference (vocoder/models/base.py:13-15) ---------------------------------------------
RR_RNN_DIMS = 256      # config/hparams.py:363
RR_FC_DIMS = 256       # :364
GN_DIMS = dict(rnn_dims=256, fc_dims=128, compute_dims=64, res_out_dims=64, res_blocks=3, upsample=(4, 5, 10), pad=2, feat=80)   # :288-300


def make_state_dict_rr(seed=0, bits=9, mode="RAW"):
    """Deterministic weights of the runtimeracer layout: the front end of oracle/weights.py (same UpsampleNetwork, res_out_dims
    128) plus I, rnn1..rnn4, fc1..fc5 of runtimeracer_version.py:119-131."""
    base = make_state_dict(seed=seed, bits=bits, mode=mode)
    sd = {k: v for k, v in base.items() if k.startswith("upsample.") or k == "step"}
    rng = np.random.default_rng(seed + 7919)

    def U(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    n_in = FEAT_DIMS + AUX_DIMS - 1 + 1
    sd["I.weight"], sd["I.bias"] = U((RR_RNN_DIMS, n_in), n_in), U((RR_RNN_DIMS,), n_in)
    for name, n_inp in (("rnn1", RR_RNN_DIMS), ("rnn2", RR_RNN_DIMS), ("rnn3", RR_RNN_DIMS + AUX_DIMS), ("rnn4", RR_RNN_DIMS)):
        sd[name + ".weight_ih_l0"] = U((3 * RR_RNN_DIMS, n_inp), RR_RNN_DIMS)
        sd[name + ".weight_hh_l0"] = U((3 * RR_RNN_DIMS, RR_RNN_DIMS), RR_RNN_DIMS)
        sd[name + ".bias_ih_l0"] = U((3 * RR_RNN_DIMS,), RR_RNN_DIMS)
        sd[name + ".bias_hh_l0"] = U((3 * RR_RNN_DIMS,), RR_RNN_DIMS)
    for name, n_inp, n_out in (("fc1", RR_RNN_DIMS + AUX_DIMS, RR_FC_DIMS), ("fc2", RR_FC_DIMS, RR_FC_DIMS), ("fc3", RR_RNN_DIMS + AUX_DIMS, RR_FC_DIMS),
                               ("fc4", RR_FC_DIMS, RR_FC_DIMS), ("fc5", RR_FC_DIMS, n_classes(bits, mode))):
        sd[name + ".weight"], sd[name + ".bias"] = U((n_out, n_inp), n_inp), U((n_out,), n_inp)
    return sd



def make_state_dict_gn(seed=0, bits=9, mode="BITS"):
    rng = np.random.default_rng(seed)

    def U(shape, fan):
        b = 1.0 / np.sqrt(fan)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    cd, ro, aux = GN_DIMS["compute_dims"], GN_DIMS["res_out_dims"], GN_DIMS["res_out_dims"] // 2
    sd = {"step": np.zeros((1,), np.int64)}

    def bn(p):
        sd[p + ".weight"] = rng.uniform(0.5, 1.5, cd).astype(np.float32)
        sd[p + ".bias"] = (0.2 * rng.standard_normal(cd)).astype(np.float32)
        sd[p + ".running_mean"] = (0.3 * rng.standard_normal(cd)).astype(np.float32)
        sd[p + ".running_var"] = rng.uniform(0.5, 1.5, cd).astype(np.float32)
        sd[p + ".num_batches_tracked"] = np.zeros((), np.int64)

    k = 2 * GN_DIMS["pad"] + 1
    sd["upsample.resnet.conv_in.weight"] = U((cd, GN_DIMS["feat"], k), GN_DIMS["feat"] * k)
    bn("upsample.resnet.batch_norm")
    for i in range(GN_DIMS["res_blocks"]):
        p = "upsample.resnet.layers.%d" % i
        sd[p + ".conv1.weight"], sd[p + ".conv2.weight"] = U((cd, cd, 1), cd), U((cd, cd, 1), cd)
        bn(p + ".batch_norm1"); bn(p + ".batch_norm2")
    sd["upsample.resnet.conv_out.weight"], sd["upsample.resnet.conv_out.bias"] = U((ro, cd, 1), cd), U((ro,), cd)
    for idx, s in zip((1, 3, 5), GN_DIMS["upsample"]):
        w = np.full((1, 1, 1, 2 * s + 1), 1.0 / (2 * s + 1), np.float64) + 0.02 * rng.standard_normal((1, 1, 1, 2 * s + 1))
        sd["upsample.up_layers.%d.weight" % idx] = w.astype(np.float32)
    C = 30 if mode == "MOL" else 2 ** bits
    n_in = GN_DIMS["feat"] + aux - 1 + 1
    R, Fc = GN_DIMS["rnn_dims"], GN_DIMS["fc_dims"]
    sd["I.weight"], sd["I.bias"] = U((R, n_in), n_in), U((R,), n_in)
    sd["rnn1.weight_ih_l0"], sd["rnn1.weight_hh_l0"] = U((3 * R, R), R), U((3 * R, R), R)
    sd["rnn1.bias_ih_l0"], sd["rnn1.bias_hh_l0"] = U((3 * R,), R), U((3 * R,), R)
    sd["fc1.weight"], sd["fc1.bias"] = U((Fc, R + aux), R + aux), U((Fc,), R + aux)
    sd["fc3.weight"], sd["fc3.bias"] = U((C, Fc), Fc), U((C,), Fc)
    return sd
