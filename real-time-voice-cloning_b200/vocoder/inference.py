"""Vocoder facade: drop-in for the reference's vocoder/inference.py (load_model :11-53, is_loaded :56-57,
infer_waveform :59-95, set_seed :97-101) with the B200 engine behind it.  Same names, positional order,
defaults, return dtype (np.float64, length (T-1)*hop) and exception types/messages.  `voc_type='libwavernn'` loads the
exported `.bin` (libwavernn_bin.py) onto the same engine."""
import numpy as np

from .. import _native
from ..config.hparams import sp, wavernn_fatchord, wavernn_geneing, wavernn_runtimeracer
from .models import base

SHARD_MIN_FOLDS = 256    # infer_waveform splits one utterance over the loaded engines only above this many folds

_model = None        # list of per-GPU WaveRNN objects once loaded
_model_type = None
_seed = 0
_calls = 0


def _devices(devices):
    if devices is None:
        return [0]
    return [int(str(d).split(":")[-1]) if not isinstance(d, int) else d for d in devices]


def load_state(state_dict, model_type=base.MODEL_TYPE_FATCHORD, devices=None, override_hp_fatchord=None, verbose=False,
               override_hp_runtimeracer=None, override_hp_geneing=None):
    """Builds the engine(s) from an in-memory state_dict (what load_model does after torch.load)."""
    global _model, _model_type
    models = []
    for d in _devices(devices):
        m, _ = base.init_voc_model(model_type, d, override_hp_fatchord=override_hp_fatchord,
                                   override_hp_runtimeracer=override_hp_runtimeracer, override_hp_geneing=override_hp_geneing)
        m.eval()
        m.load_state_dict(state_dict)
        m.precision = _native.PREC_AUTO       # facade default: the fastest loop for the call (fatchord_version.resolve_precision)
        models.append(m)
    _model, _model_type = models, model_type
    if verbose:
        print("Model has been trained to step %d." % models[0].get_step())
    return models[0]


def load_model(weights_fpath, voc_type=base.VOC_TYPE_PYTORCH, verbose=True, devices=None):
    """inference.py:11.  voc_type 'pytorch' and 'b200' both select this engine (it replaces the PyTorch
    path); 'libwavernn' decodes the exported .bin (fatchord topology) onto the same engine."""
    global _model, _model_type
    if voc_type in (base.VOC_TYPE_PYTORCH, base.VOC_TYPE_B200):
        import torch  # checkpoint I/O only
        checkpoint = torch.load(weights_fpath, map_location="cpu")
        model_type = base.MODEL_TYPE_FATCHORD
        if "model_type" in checkpoint:
            model_type = checkpoint["model_type"]
        state = checkpoint["model_state"] if "model_state" in checkpoint else checkpoint   # legacy bare dict, :417-424
        try:
            load_state(state, model_type, devices=devices)
        except NotImplementedError as e:      # inference.py:27-32 prints and returns
            print(str(e))
            return
        if verbose:
            print("Loaded vocoder of model '%s' at path '%s'." % (_model_type, weights_fpath))
            print("Model has been trained to step %d." % (_model[0].get_step()))
    elif voc_type == base.VOC_TYPE_CPP:
        # inference.py:41-50 hands the path to the C++ vocoder; here the exported (usually pruned) weights are decoded
        # and run on the B200 engine -- pruned checkpoints take the block-sparse loop.  The .bin carries no hparams:
        # mode / bits follow from fc3's row count (30 -> MOL, 2**bits -> RAW), everything else must match the fatchord
        # hparams the engine is built for.  The reference hard-codes the runtimeracer topology on this path
        # (inference.py:43, FIXME there); that topology is SURVEY.md section 8(f) "next".
        import copy
        from . import libwavernn_bin
        try:
            state, meta = libwavernn_bin.read_bin(weights_fpath)
        except OSError:
            raise RuntimeError("Cannot open file.")                       # WaveRNNVocoder.cpp:24-26
        hp = copy.deepcopy(wavernn_fatchord)
        if (meta["res_blocks"], tuple(meta["upsample_factors"]), meta["pad"]) != (hp.res_blocks, tuple(hp.upsample_factors), hp.pad):
            raise NotImplementedError("libwavernn file was exported with other hparams than the fatchord vocoder "
                                      "(res_blocks %d, upsample %s, pad %d)" % (meta["res_blocks"], meta["upsample_factors"], meta["pad"]))
        C = meta["n_classes"]
        if C == 30:
            hp.mode = "MOL"
        elif C >= 2 and (C & (C - 1)) == 0:
            hp.mode, hp.bits = "RAW", C.bit_length() - 1
        else:
            raise NotImplementedError("libwavernn file has %d output classes: neither MOL (30) nor RAW (2**bits)" % C)
        load_state(state, base.MODEL_TYPE_FATCHORD, devices=devices, override_hp_fatchord=hp)
        if verbose:
            print("Loaded vocoder of model '%s' at path '%s'." % (_model_type, weights_fpath))
    else:
        raise NotImplementedError("Invalid vocoder of type '%s' provided. Aborting..." % voc_type)


def unload():
    global _model, _model_type
    _model, _model_type = None, None


def is_loaded():
    return _model is not None


def set_seed(seed):
    """inference.py:97: the reference seeds torch's global RNG; here the seed keys the Philox stream."""
    global _seed, _calls
    _seed, _calls = int(seed), 0


def _next_seed():
    global _calls
    s = (_seed + 0x9E3779B97F4A7C15 * _calls) & 0xFFFFFFFFFFFFFFFF   # successive calls draw fresh noise
    _calls += 1
    return s


def _hp():
    if _model_type == base.MODEL_TYPE_FATCHORD:
        return wavernn_fatchord
    if _model_type == base.MODEL_TYPE_RUNTIMERACER:               # inference.py:69-70
        return wavernn_runtimeracer
    if _model_type == base.MODEL_TYPE_GENEING:                    # inference.py:67-68
        return wavernn_geneing
    raise NotImplementedError("Invalid model of type '%s' provided. Aborting..." % _model_type)


def infer_waveform(mel, normalize=True, batched=True, target=None, overlap=None, progress_callback=None):
    """inference.py:59.  mel: (80, T) float32 in the synthesizer's range.  With several engines loaded
    (load_model(..., devices=[...])) the folds of the utterance are split into contiguous ranges, one
    per GPU, run concurrently from host threads, gathered on the host and cross-faded on GPU 0 -- no
    collective is involved."""
    if _model is None or _model_type is None:
        raise Exception("Please load Wave-RNN in memory before using it")
    hp_wavernn = _hp()
    if target is None:
        target = hp_wavernn.gen_target
    if overlap is None:
        overlap = hp_wavernn.gen_overlap
    if normalize:
        mel = mel / sp.max_abs_value
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    seed = _next_seed()
    shard = len(_model) > 1 and batched
    if shard:
        # Splitting ONE utterance pays only where the loop's time grows with the fold count.  Below ~256 folds a step is a latency
        # chain whose length does not depend on the work (10.6 us at 18 folds, 12.7 us at 213), so two GPUs
        # with half the folds each finish hardly earlier and the host gather eats the rest (measured on 2 B200s, cfg3ref: 73 ms
        # sharded vs 77 ms on one GPU).
        try:
            shard = _native.fold_plan(mel.shape[1] * sp.hop_size, int(target), int(overlap))[0] > SHARD_MIN_FOLDS
        except Exception:
            shard = False
    if not shard:
        return _model[0].generate(mel[None, ...], batched, target, overlap, hp_wavernn.mu_law, sp.preemphasize, progress_callback,
                                  seed=seed)
    return _infer_sharded(mel, target, overlap, hp_wavernn.mu_law, sp.preemphasize, seed, progress_callback)


def _infer_sharded(mel, target, overlap, mu_law, preemph, seed, progress_callback=None):
    """Fold ranges of ONE utterance on several engines.  The precision is chosen once from the utterance's total fold count
    (not per shard), the shards write straight into one (F, S) sample matrix, and progress_callback keeps the reference's
    contract (fatchord_version.py:234-236: (i, seq_len, b_size, gen_rate_kHz)) with b_size = all folds of the utterance and
    i = the slowest shard's step."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    from .. import _native
    from .models.fatchord_version import resolve_precision
    T = mel.shape[1]
    F, _ = _native.fold_plan(T * sp.hop_size, target, overlap)
    S = target + 2 * overlap
    n = min(len(_model), F)
    bounds = [F * i // n for i in range(n + 1)]
    m0 = _model[0]
    prec = m0.precision
    if prec == _native.PREC_AUTO:
        prec = resolve_precision(m0.n_classes, m0.sparsity, F, m0.sparse_available)
    samples = np.empty((F, S), np.float32)
    steps, lock = [0] * n, threading.Lock()

    def work(i):
        cb = None
        if progress_callback is not None:
            def cb(step, seq_len, b_size, rate, i=i):
                with lock:
                    steps[i] = step
                    progress_callback(min(steps), seq_len, F, rate * F / max(1, b_size))
        samples[bounds[i]:bounds[i + 1]] = _model[i].generate_debug(mel, True, target, overlap, fold_begin=bounds[i], fold_end=bounds[i + 1],
                                                                    seed=seed, precision=prec, progress_callback=cb)["samples"]
    with ThreadPoolExecutor(max_workers=n) as ex:
        list(ex.map(work, range(n)))
    return m0.postprocess(samples, True, overlap, T, mu_law, preemph)


def infer_waveforms(mels, normalize=True, batched=True, target=None, overlap=None, utt_index0=0, progress_callback=None):
    """Many utterances: sharded by utterance across the loaded engines (BASELINE config 5); inside one
    engine all folds share persistent-loop launches.  Returns a list of float64 arrays.  The loop precision is chosen ONCE
    from the corpus' total fold count, so the result does not depend on how it is sharded; progress_callback (if given) is
    called with the reference's four arguments for every shard's launches."""
    if _model is None or _model_type is None:
        raise Exception("Please load Wave-RNN in memory before using it")
    from concurrent.futures import ThreadPoolExecutor
    hp_wavernn = _hp()
    target = hp_wavernn.gen_target if target is None else target
    overlap = hp_wavernn.gen_overlap if overlap is None else overlap
    mels = [np.ascontiguousarray(m / sp.max_abs_value if normalize else m, dtype=np.float32) for m in mels]
    seed = _next_seed()
    from .models.fatchord_version import resolve_precision
    m0 = _model[0]
    prec = m0.precision
    if prec == _native.PREC_AUTO:
        prec = resolve_precision(m0.n_classes, m0.sparsity, m0._count_folds(mels, batched, target, overlap), m0.sparse_available)
    n = min(len(_model), len(mels))
    # balance by total frames: longest first onto the lightest shard
    order = sorted(range(len(mels)), key=lambda i: -mels[i].shape[1])
    shards, load = [[] for _ in range(n)], [0] * n
    for i in order:
        k = load.index(min(load))
        shards[k].append(i)
        load[k] += mels[i].shape[1]
    out = [None] * len(mels)

    def work(k):
        m = _model[k]
        idx = sorted(shards[k])
        # contiguous runs share a call; non-contiguous indices need their own utt_index0
        runs, start = [], 0
        for j in range(1, len(idx) + 1):
            if j == len(idx) or idx[j] != idx[j - 1] + 1:
                runs.append(idx[start:j])
                start = j
        for run in runs:
            wavs = m.generate_batch([mels[i] for i in run], batched, target, overlap, hp_wavernn.mu_law,
                                    sp.preemphasize, progress_callback, utt_index0=utt_index0 + run[0], seed=seed, precision=prec)
            for i, w in zip(run, wavs):
                out[i] = w
    with ThreadPoolExecutor(max_workers=n) as ex:
        list(ex.map(work, range(n)))
    return out
