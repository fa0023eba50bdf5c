"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11) in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The CUDA sample loop draws its noise from the
same generator (csrc/philox.cuh) so that a run can be replayed on the reference: the reference's own
sampling rule (torch Categorical = exponential race; uniform_() for MOL,
vocoder/distribution.py:123,135) is replaced by "one uniform per (step, fold)" as BASELINE.json's
north_star requires ("counter-based Philox stream that can be replayed on the reference").

Noise contract (shared with csrc/philox.cuh):
  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (step, fold_in_utterance, utterance_index, block)
  uniform  = ((x >> 8) + 0.5) * 2**-24          -> strictly inside (0, 1), exactly representable in fp32
  RAW      : u  = uniform(word 0 of block 0)
  MOL      : mixture uniforms j=0..9 = word j%4 of block j//4 ; logistic uniform = word 2 of block 2
             both mapped to [1e-5, 1-1e-5] as  1e-5 + u * (1 - 2e-5)  in fp32 (distribution.py:123,135)
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0
            p1 = _M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
            n0 = hi1 ^ c1 ^ np.uint64(k0)
            n2 = hi0 ^ c3 ^ np.uint64(k1)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    """uint32 -> float32 uniform strictly inside (0,1)."""
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def raw_uniforms(seed, steps, folds, utt=0, fold0=0):
    """(steps, folds) float32 uniforms for the RAW inverse-CDF sampler."""
    k0, k1 = _key(seed)
    s = np.arange(steps, dtype=np.uint64)[:, None]
    f = (np.arange(folds, dtype=np.uint64) + np.uint64(fold0))[None, :]
    x0, _, _, _ = philox4x32_10(s, f, utt, 0, k0, k1)
    return u01(x0)


def mol_uniforms(seed, steps, folds, utt=0, fold0=0):
    """Returns (u_mix (steps, folds, 10), u_logistic (steps, folds)), float32 in [1e-5, 1-1e-5]."""
    k0, k1 = _key(seed)
    s = np.arange(steps, dtype=np.uint64)[:, None]
    f = (np.arange(folds, dtype=np.uint64) + np.uint64(fold0))[None, :]
    blocks = [philox4x32_10(s, f, utt, b, k0, k1) for b in range(3)]
    words = [u01(blocks[j // 4][j % 4]) for j in range(11)]
    lo = np.float32(1e-5)
    span = np.float32(1.0) - np.float32(2e-5)
    words = [lo + w * span for w in words]
    return np.stack(words[:10], axis=-1), words[10]
