"""WaveRNN (fatchord topology) backed by the B200 engine.

Mirrors the inference surface of the reference class (vocoder/models/fatchord_version.py:88-436):
generate (:155), fold_with_overlap (:290), xfade_and_unfold (:342), pad_tensor (:275), get_step (:406),
load_state_dict / state_dict / eval / train as used by vocoder/inference.py:11-53.  The object owns a native
engine handle (include/wavernn_b200.h); all arithmetic runs in CUDA on the engine's GPU.

Deliberate differences from the reference (DESIGN.md "quirks"):
  * Q1: generate() does not flip the model to train mode (there is no train mode here).
  * Q3: no torch RNG is consumed; noise comes from Philox(seed, step, fold, utterance).
  * Q4: RAW sampling is inverse-CDF on one uniform per (step, fold) (north_star's replayable rule).
"""
import ctypes as C
import threading
import time

import numpy as np

from ... import _native

_FIXED = dict(rnn_dims=512, fc_dims=512, pad=2, upsample_factors=(5, 5, 8), feat_dims=80, compute_dims=128,
              res_out_dims=128, res_blocks=10, hop_length=200)


def _device_index(device):
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    idx = getattr(device, "index", None)          # torch.device
    if idx is not None:
        return int(idx)
    s = str(device)
    return int(s.split(":")[1]) if ":" in s else 0


def _raise(lib, handle, rc):
    msg = lib.wrnn_last_error(handle)
    msg = msg.decode() if msg else "error %d" % rc
    if rc == _native.ERR_NOT_LOADED:
        raise Exception(msg)
    if rc in (_native.ERR_TOO_SHORT, _native.ERR_INVALID, _native.ERR_SHAPE):
        raise ValueError(msg)
    raise RuntimeError(msg)


class WaveRNN(object):
    _FIXED_DIMS = _FIXED                      # the topology this class is built for (subclasses: runtimeracer_version.py)
    _TOPOLOGY = _native.TOPO_FATCHORD

    def __init__(self, rnn_dims, fc_dims, bits, pad, upsample_factors, feat_dims, compute_dims, res_out_dims,
                 res_blocks, hop_length, sample_rate, mode='RAW', pruning=False, device=0):
        given = dict(rnn_dims=rnn_dims, fc_dims=fc_dims, pad=pad, upsample_factors=tuple(upsample_factors),
                     feat_dims=feat_dims, compute_dims=compute_dims, res_out_dims=res_out_dims,
                     res_blocks=res_blocks, hop_length=hop_length)
        for k, v in self._FIXED_DIMS.items():
            if given[k] != v:
                raise NotImplementedError("the B200 engine is built for %s=%r (got %r)" % (k, v, given[k]))
        self.mode = mode
        if mode == 'RAW':
            self.n_classes = 2 ** bits
        elif mode == 'MOL':
            self.n_classes = 30
        else:
            raise RuntimeError("Unknown model mode value - ", mode)   # fatchord_version.py:100,232
        self.bits = bits
        self.pad = pad
        self.rnn_dims = rnn_dims
        self.aux_dims = res_out_dims // 4
        self.hop_length = hop_length
        self.sample_rate = sample_rate
        self.device_index = _device_index(device)
        self.precision = _native.PREC_F32
        self.seed = 0
        self.last_timings = {}
        self._state = {}
        self._lib = _native.load()
        self._h = C.c_void_p()
        rc = self._lib.wrnn_create(self.device_index, bits, _native.MODE_RAW if mode == 'RAW' else _native.MODE_MOL,
                                   C.byref(self._h))
        if rc == _native.ERR_INVALID:
            raise ValueError("wrnn_create: unsupported bits=%r / mode=%r (RAW takes 8..10 bits)" % (bits, mode))
        if rc != _native.OK:
            raise RuntimeError("wrnn_create failed (%d): no usable CUDA device %d" % (rc, self.device_index))
        if self._TOPOLOGY != _native.TOPO_FATCHORD and self._lib.wrnn_set_topology(self._h, self._TOPOLOGY) != _native.OK:
            raise RuntimeError("wrnn_set_topology(%d) failed" % self._TOPOLOGY)
        self._lock = threading.Lock()        # a handle is not re-entrant (include/wavernn_b200.h): calls on it are serialised here

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.wrnn_destroy(h)
            self._h = C.c_void_p()

    # ---- checkpoint I/O (vocoder/inference.py:21-36, train.py:316-324 layout) ---------------------------
    def load_state_dict(self, state_dict, strict=True):
        """Accepts torch tensors or numpy arrays keyed like the reference state_dict.  A pruned checkpoint
        is the same dense dict with zeros (vocoder/pruner.py:55-58); its block pattern is re-derived.
        strict=True (torch's default): a missing tensor is a RuntimeError (the engine reports which one)."""
        state = {}
        for name, value in state_dict.items():
            if hasattr(value, "detach"):
                value = value.detach().cpu().numpy()
            state[name] = np.asarray(value)
        for name, arr in state.items():
            if name == "step":
                self._lib.wrnn_set_step(self._h, int(arr.reshape(-1)[0]))
                continue
            if arr.dtype.kind in "iu":            # num_batches_tracked
                continue
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
            rc = self._lib.wrnn_set_tensor(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim)
            if rc != _native.OK:
                _raise(self._lib, self._h, rc)
        rc = self._lib.wrnn_finalize(self._h)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        self._state = state
        return self

    def state_dict(self):
        return dict(self._state)

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def to(self, device):
        return self

    def get_step(self):
        return int(self._lib.wrnn_get_step(self._h))

    def num_params(self, print_out=True):
        n = sum(int(np.prod(v.shape)) for k, v in self._state.items()
                if v.dtype.kind == "f" and "running_" not in k) / 1_000_000
        if print_out:
            print('Trainable Parameters: %.3fM' % n)
        return n

    @property
    def sparsity(self):
        return float(self._lib.wrnn_sparsity(self._h))

    @property
    def sparse_available(self):
        """The block-sparse loop can run this checkpoint (its compressed images fit one cluster's shared memory)."""
        return bool(self._lib.wrnn_sparse_available(self._h))

    @property
    def launch_count(self):
        return int(self._lib.wrnn_launch_count(self._h))

    def _count_folds(self, arrs, batched, target, overlap):
        if not batched:
            return len(arrs)
        n = 0
        for a in arrs:
            try:
                n += _native.fold_plan(a.shape[1] * self.hop_length, int(target), int(overlap))[0]
            except Exception:
                n += 1                            # (the engine reports the bad plan with the reference's message)
        return n

    # ---- the hot path ---------------------------------------------------------------------------------------
    def _request(self, mels, batched, target, overlap, mu_law, apply_preemphasis, progress_callback, want_wav=True,
                 **extra):
        arrs, ptrs, keep = [], [], []
        for m in mels:
            if hasattr(m, "detach"):
                m = m.detach().cpu().numpy()
            m = np.asarray(m)
            if m.ndim == 3:                      # (1, 80, T) as infer_waveform passes it, inference.py:93
                if m.shape[0] != 1:
                    raise ValueError("generate() takes one utterance: mels.size(0) must be 1")
                m = m[0]
            m = np.ascontiguousarray(m, dtype=np.float32)
            if m.ndim != 2 or m.shape[0] != 80:
                raise ValueError("mel must be (80, T) float32")
            arrs.append(m)
        n = len(arrs)
        rq = _native.Request()
        rq.n_utts = n
        mel_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        Ts = (C.c_int32 * n)(*[a.shape[1] for a in arrs])
        rq.mels = C.cast(mel_ptrs, C.POINTER(C.c_void_p))
        rq.T = C.cast(Ts, C.POINTER(C.c_int32))
        rq.batched = 1 if batched else 0
        rq.target = int(target) if target is not None else 0
        rq.overlap = int(overlap) if overlap is not None else 0
        rq.mu_law = 1 if mu_law else 0
        rq.apply_preemphasis = 1 if apply_preemphasis else 0
        prec = int(extra.get("precision", self.precision))
        if self._TOPOLOGY != _native.TOPO_FATCHORD:
            prec = _native.PREC_F32              # (the other topologies have the fp32 loop only)
        elif prec == _native.PREC_AUTO:
            prec = resolve_precision(self.n_classes, self.sparsity, self._count_folds(arrs, batched, target, overlap),
                                     self.sparse_available)
        rq.precision = prec
        rq.seed = int(extra.get("seed", self.seed)) & 0xFFFFFFFFFFFFFFFF
        rq.utt_index0 = int(extra.get("utt_index0", 0))
        rq.fold_begin = int(extra.get("fold_begin", 0))
        rq.fold_end = int(extra.get("fold_end", 0))
        rq.max_steps = int(extra.get("max_steps", 0))
        keep += [arrs, mel_ptrs, Ts]
        if progress_callback is not None:
            cb = _native.PROGRESS_FN(lambda i, s, b, r, u: progress_callback(int(i), int(s), int(b), float(r)))
            rq.progress = cb
            keep.append(cb)
        offsets = (C.c_int64 * (n + 1))()
        rq.wav_offsets = C.cast(offsets, C.POINTER(C.c_int64))
        wav = None
        if want_wav:
            total = sum(max(0, (a.shape[1] - 1) * self.hop_length) for a in arrs)
            wav = np.empty(max(1, total), np.float64)
            rq.wav = wav.ctypes.data
            rq.wav_capacity = total
        keep.append(offsets)
        return rq, arrs, wav, offsets, keep

    def _run(self, rq):
        with self._lock:
            rc = self._lib.wrnn_generate(self._h, C.byref(rq))
            if rc != _native.OK:
                _raise(self._lib, self._h, rc)
        self.last_timings = dict(ms_h2d=rq.ms_h2d, ms_cond=rq.ms_cond, ms_loop=rq.ms_loop, ms_post=rq.ms_post,
                                 ms_d2h=rq.ms_d2h, n_folds=rq.n_folds, n_steps=rq.n_steps, n_launches=rq.n_launches,
                                 precision=int(rq.precision), loop_kernel=_native.LOOP_KERNELS.get(int(rq.loop_kernel), "?"))

    def generate(self, mels, batched, target, overlap, mu_law, apply_preemphasis, progress_callback=None, seed=None):
        """fatchord_version.py:155: mels is (1, 80, T) float32 already divided by max_abs_value; returns
        np.float64[(T-1)*hop].  Raises ValueError for T <= 20 like the reference's broadcast error (Q8).
        `seed` (extension): Philox key of this call; default self.seed."""
        extra = {} if seed is None else dict(seed=seed)
        rq, arrs, wav, offsets, keep = self._request([mels], batched, target, overlap, mu_law, apply_preemphasis,
                                                     progress_callback, **extra)
        self._run(rq)
        return wav[:(arrs[0].shape[1] - 1) * self.hop_length]

    def generate_batch(self, mels_list, batched, target, overlap, mu_law, apply_preemphasis, progress_callback=None,
                       utt_index0=0, seed=None, precision=None):
        """Many utterances in one call: their folds are pooled into the same persistent-loop launches
        (BASELINE config 5).  Returns a list of float64 arrays.  Utterance i uses Philox utterance
        counter utt_index0 + i, so the result does not depend on how a corpus is sharded."""
        extra = dict(utt_index0=utt_index0)
        if seed is not None:
            extra["seed"] = seed
        if precision is not None:
            extra["precision"] = precision
        rq, arrs, wav, offsets, keep = self._request(mels_list, batched, target, overlap, mu_law, apply_preemphasis,
                                                     progress_callback, **extra)
        self._run(rq)
        return [wav[offsets[i]:offsets[i + 1]].copy() for i in range(len(arrs))]

    def generate_debug(self, mels, batched, target, overlap, forced=None, max_steps=0, want_logits=False,
                       fold_begin=0, fold_end=0, seed=None, utt_index0=0, precision=None, progress_callback=None):
        """Parity hook: runs the loop and returns dict(samples (F,S), logits (F,S,C) or None).  `forced`
        (F,S) float32 replaces the fed-back samples (teacher forcing on the reference's samples)."""
        extra = dict(max_steps=max_steps, fold_begin=fold_begin, fold_end=fold_end, utt_index0=utt_index0)
        if seed is not None:
            extra["seed"] = seed
        if precision is not None:
            extra["precision"] = precision
        rq, arrs, _, _, keep = self._request([mels], batched, target, overlap, True, True, progress_callback, want_wav=False, **extra)
        T = arrs[0].shape[1]
        N = T * self.hop_length
        if batched:
            F, _ = _native.fold_plan(N, target, overlap)
            S = target + 2 * overlap
        else:
            F, S = 1, N
        if fold_begin or fold_end:
            F = min(F, fold_end) - max(0, fold_begin)
        S_run = min(S, max_steps) if max_steps else S
        if forced is not None:
            forced = np.ascontiguousarray(forced, dtype=np.float32)
            assert forced.shape == (F, S), (forced.shape, (F, S))
            rq.forced = forced.ctypes.data
        samples = np.zeros((F, S_run), np.float32)
        rq.samples = samples.ctypes.data
        logits = None
        if want_logits:
            logits = np.zeros((F, S_run, self.n_classes), np.float32)
            rq.logits = logits.ctypes.data
        self._run(rq)
        return dict(samples=samples, logits=logits)

    # ---- helpers callers may touch -----------------------------------------------
    def conditioning(self, mel):
        """UpsampleNetwork.forward (fatchord_version.py:78-85) for one normalised (80,T) mel: returns
        (mels_up (200T, 80) rebuilt from the engine's interpolation table, aux frames (T, 128))."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        T = mel.shape[1]
        aux = np.zeros((T, 128), np.float32)
        up = np.zeros((T * self.hop_length, 80), np.float32)
        rc = self._lib.wrnn_condition(self._h, mel.ctypes.data, T, aux.ctypes.data, up.ctypes.data)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return up, aux

    def conditioning_tc(self, mel):
        """MelResNet output (T,128) computed by the tensor-core front end (cond_tc.cu)."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        aux = np.zeros((mel.shape[1], 128), np.float32)
        rc = self._lib.wrnn_condition_tc(self._h, mel.ctypes.data, mel.shape[1], aux.ctypes.data)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return aux

    def pad_tensor(self, x, pad, side='both'):
        """fatchord_version.py:275-288 on a (b, t, c) torch tensor (same device as x)."""
        import torch
        b, t, c = x.size()
        total = t + 2 * pad if side == 'both' else t + pad
        padded = torch.zeros(b, total, c, dtype=x.dtype, device=x.device)
        if side == 'before' or side == 'both':
            padded[:, pad:pad + t, :] = x
        elif side == 'after':
            padded[:, :t, :] = x
        return padded

    def fold_with_overlap(self, x, target, overlap):
        """fatchord_version.py:290-340 on a (1, total_len, features) torch tensor.  Pure indexing (the
        engine itself never materialises folds: its loop indexes per-frame tables by fold offset); the
        plan comes from the same native routine the engine uses (wrnn_fold_plan)."""
        import torch
        _, total_len, features = x.size()
        num_folds, padded_len = _native.fold_plan(total_len, target, overlap)
        if padded_len != total_len:
            x = self.pad_tensor(x, padded_len - total_len, side='after')
        S = target + 2 * overlap
        idx = (torch.arange(num_folds, device=x.device) * (target + overlap))[:, None] + torch.arange(S, device=x.device)[None]
        return x[0][idx.reshape(-1)].reshape(num_folds, S, features).contiguous()

    def xfade_and_unfold(self, y, target, overlap):
        """fatchord_version.py:342-404 on the GPU; y (num_folds, length) float64 ndarray.  `target` is
        ignored exactly like the reference does (Q7).  Unlike the reference, y is not modified."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        F, S = y.shape
        out = np.empty(F * (S - overlap) + overlap, np.float64)
        rc = self._lib.wrnn_xfade_unfold(self._h, y.ctypes.data, F, S, int(overlap), out.ctypes.data)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return out

    def postprocess(self, samples, batched, overlap, T, mu_law, apply_preemphasis):
        """Tail of generate (fatchord_version.py:238-255) on host (F,S) float32 samples -- used when folds
        of one utterance were generated on several GPUs and gathered on the host."""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        F, S = samples.shape
        wav = np.empty((T - 1) * self.hop_length, np.float64)
        with self._lock:
            rc = self._lib.wrnn_postprocess(self._h, samples.ctypes.data, F, S, 1 if batched else 0, int(overlap), int(T),
                                            1 if mu_law else 0, 1 if apply_preemphasis else 0, wav.ctypes.data)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return wav

    def barrier_floor(self, rounds=20000):
        ll, cnt = C.c_float(), C.c_float()
        rc = self._lib.wrnn_barrier_floor(self._h, rounds, C.byref(ll), C.byref(cnt))
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return dict(ll_us=ll.value, counter_us=cnt.value)

    def cluster_floor(self, cluster_size=16, rounds=20000):
        us = C.c_float()
        rc = self._lib.wrnn_cluster_floor(self._h, cluster_size, rounds, C.byref(us))
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return us.value

    def debug_tc_gemm(self, A, W):
        """Self-test of the tcgen05/TMA building blocks: A (128,512), W (N,512) float16 -> (128,N) float32."""
        A = np.ascontiguousarray(A, dtype=np.float16)
        W = np.ascontiguousarray(W, dtype=np.float16)
        assert A.shape == (128, 512) and W.shape[1] == 512
        out = np.zeros((128, W.shape[0]), np.float32)
        rc = self._lib.wrnn_debug_tc_gemm(self._h, A.ctypes.data, W.ctypes.data, W.shape[0], out.ctypes.data)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return out

    def debug_tc_gemm2(self, A, W):
        """Self-test of the CTA-pair path (tcgen05 cta_group::2): A (256,512), W (N,512) float16 -> (256,N) float32."""
        A = np.ascontiguousarray(A, dtype=np.float16)
        W = np.ascontiguousarray(W, dtype=np.float16)
        assert A.shape == (256, 512) and W.shape[1] == 512
        out = np.zeros((256, W.shape[0]), np.float32)
        rc = self._lib.wrnn_debug_tc_gemm2(self._h, A.ctypes.data, W.ctypes.data, W.shape[0], out.ctypes.data)
        if rc != _native.OK:
            _raise(self._lib, self._h, rc)
        return out

    def gen_display(self, i, seq_len, b_size, gen_rate):
        """Default progress line (fatchord_version.py:262-265)."""
        done = int(16 * i // max(1, seq_len))
        bar = '#' * done + '-' * (16 - done)
        print('\r| %s %d/%d | Batch Size: %d | Gen Rate: %.1fkHz | ' % (bar, i * b_size, seq_len * b_size, b_size,
                                                                       gen_rate), end='', flush=True)


AUTO_F16_MIN_FOLDS = 8
AUTO_F16_MIN_FOLDS_MOL = 4
AUTO_SPARSE_MIN = 0.8


def resolve_precision(n_classes, sparsity, n_folds, sparse_available=True):
    """Which loop `PREC_AUTO` (the default of the `vocoder.inference` facade) runs, from what was measured on B200 (DESIGN.md
    sections 4.5, 6): a pruned checkpoint whose compressed images fit one cluster -> the block-sparse cluster loop (62.9x vs
    38.8x real-time on cfg4; a pruned checkpoint that does NOT fit falls through to the dense loops -- its tensors are dense
    with zeros, vocoder/pruner.py:55-58); a dense one with at least AUTO_F16_MIN_FOLDS folds in the call -> the fp16
    tensor-core loops (MOL up to 384 folds, RAW up to 256: the role-specialised loop_rs.cu; MOL 385..768: two waves of it; above: loop_tc.cu); fewer folds -> the
    fp32 loop, which is then as fast AND bit-faithful (one fold: the only loop built for it).  Measured crossover, us per
    step fp32 loop / loop_rs.cu (tools/auto_crossover_raw.py, RAW 9-bit): 2 folds 14.6 / 14.3, 5 folds 16.1 / 15.5, 9 folds
    18.6 / 15.7, 13 folds 20.5 / 15.9, 19 folds 26.2 / 16.1; MOL on loop_rs.cu runs 10.6 us per step at 18 folds.
    The tensor-core loops exist for 30 (MOL), 512 and 1024 classes."""
    if sparsity >= AUTO_SPARSE_MIN and sparse_available:
        return _native.PREC_SPARSE_F32
    min_folds = AUTO_F16_MIN_FOLDS_MOL if n_classes == 30 else AUTO_F16_MIN_FOLDS
    if n_folds >= min_folds and n_classes in (30, 512, 1024):
        return _native.PREC_F16
    return _native.PREC_F32


def _now():
    return time.time()
