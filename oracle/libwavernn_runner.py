"""Python side of the libwavernn comparator (reference: vocoder/libwavernn/inference.py:37-54 one engine per core,
:56-128 vocode_mel, :135-198 fold / unfold in the mel-frame domain).  TEST / BENCH INFRASTRUCTURE ONLY."""
import ctypes as C
import math
import os
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import wavernn_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libwavernn_port.so")


def build(force=False):
    subprocess.run(["make", "-s", "-C", HERE] + (["-B"] if force else []), check=True)
    return LIB


class PortVocoder:
    def __init__(self, bin_path, threads=1):
        build()
        self.lib = C.CDLL(LIB)
        self.lib.lwr_load.restype = C.c_void_p
        self.lib.lwr_load.argtypes = [C.c_char_p]
        self.lib.lwr_free.argtypes = [C.c_void_p]
        self.lib.lwr_mel_to_wav.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.handles = [self.lib.lwr_load(str(bin_path).encode()) for _ in range(threads)]   # inference.py:47-54
        if not all(self.handles):
            raise RuntimeError("Cannot open file.")                                          # WaveRNNVocoder.cpp:24-26

    def __del__(self):
        for h in getattr(self, "handles", []):
            if h:
                self.lib.lwr_free(h)

    def mel_to_wav(self, tid, mel, uniforms=None):
        """Vocoder::melToWav (WaveRNNVocoder.cpp:37-47): (80,T) -> float32[200*T] label-floats in [-1,1]."""
        mel = np.ascontiguousarray(mel, np.float32)
        out = np.empty(mel.shape[1] * 200, np.float32)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, np.float32)
        self.lib.lwr_mel_to_wav(self.handles[tid], mel.ctypes.data, mel.shape[1], None if u is None else u.ctypes.data,
                                out.ctypes.data)
        return out

    def vocode_mel(self, mel_norm, min_target, min_overlap, n_classes, mu_law=True, preemph=True):
        """libwavernn/inference.py:56-128 (mel already normalised)."""
        hop = 200
        wave_len = mel_norm.shape[1] * hop
        n = len(self.handles)
        if n == 1:
            output = self.mel_to_wav(0, mel_norm).astype(np.float64)
        else:
            optimal_target = max(math.ceil(((wave_len - min_overlap) / n) - min_overlap), min_target)
            mt, mo = math.ceil(optimal_target / hop), math.ceil(min_overlap / hop)
            T = mel_norm.shape[1]
            folds, padded = orc.fold_plan(T, mt, mo)
            mp = np.zeros((mel_norm.shape[0], padded), np.float32)
            mp[:, :T] = mel_norm
            chunks = [mp[:, i * (mt + mo): i * (mt + mo) + mt + 2 * mo] for i in range(folds)]
            with ThreadPoolExecutor(max_workers=n) as ex:
                outs = list(ex.map(lambda a: self.mel_to_wav(a[0] % n, a[1]), enumerate(chunks)))
            output = orc.xfade_and_unfold(np.stack(outs).astype(np.float64), mo * hop)
        if mu_law:
            output = orc.decode_mu_law(output, n_classes)
        if preemph:
            output = orc.de_emphasis(output)
        output = output[:wave_len].copy()
        output[-20 * hop:] *= np.linspace(1, 0, 20 * hop)
        return output


def time_threads(bin_path, mel_norm, threads, frames_per_thread):
    """Every thread vocodes its own `frames_per_thread`-frame chunk concurrently (GIL released in the C call):
    returns (samples generated, seconds)."""
    v = PortVocoder(bin_path, threads)
    chunks = [np.ascontiguousarray(mel_norm[:, (i * 7) % max(1, mel_norm.shape[1] - frames_per_thread):][:, :frames_per_thread])
              for i in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        outs = list(ex.map(lambda a: v.mel_to_wav(a[0], a[1]), enumerate(chunks)))
    dt = time.perf_counter() - t0
    return sum(o.size for o in outs), dt
