"""The callers' spectrogram hand-off (vocoder/handoff.py) end to end on the GPU: several texts ->
one waveform with 0.15 s gaps, the reference's way (concatenate, vocode once, cut: toolbox/toolbox.py:263-321) and pooled (every
text its own utterance in one engine call)."""
import copy

import numpy as np
import pytest

from oracle import weights
from tests.util import norm_mel

pytestmark = pytest.mark.gpu


def test_handoff_on_gpu_both_ways():
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import handoff, inference
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_fatchord)
    hp.bits, hp.mode = 9, "RAW"
    inference.load_state(weights.make_state_dict(seed=11, bits=9, mode="RAW"), override_hp_fatchord=hp)
    try:
        specs = [norm_mel(T, 30 + i) * 4.0 for i, T in enumerate((31, 44, 26))]
        # (a) the reference's way: literal restatement of toolbox.py:264-265, 309-314, 321 around ONE infer_waveform call
        inference.set_seed(5)
        got = handoff.vocode_specs(specs, target=800, overlap=200)
        inference.set_seed(5)
        spec = np.concatenate(specs, axis=1)
        wav = inference.infer_waveform(spec, target=800, overlap=200)
        breaks = [s.shape[1] for s in specs]
        b_ends = np.cumsum(np.array(breaks) * 200)
        b_starts = np.concatenate(([0], b_ends[:-1]))
        wavs = [wav[a:b] for a, b in zip(b_starts, b_ends)]
        gaps = [np.zeros(int(0.15 * 16000))] * len(breaks)
        want = np.concatenate([i for w, g in zip(wavs, gaps) for i in (w, g)])
        want = want / np.abs(want).max() * 0.97
        assert got.dtype == np.float64 and got.shape == want.shape == ((sum(breaks) - 1) * 200 + 3 * 2400,)
        assert np.array_equal(got, want)
        # (b) pooled: piece i is utterance i of one infer_waveforms call, (T_i - 1) * hop samples, then the gap
        inference.set_seed(5)
        pooled = handoff.vocode_specs(specs, normalize_peak=False, pooled=True, target=800, overlap=200)
        inference.set_seed(5)
        each = inference.infer_waveforms(specs, target=800, overlap=200)
        assert [len(w) for w in each] == [(T - 1) * 200 for T in breaks]
        pos = 0
        for w in each:
            assert np.array_equal(pooled[pos:pos + len(w)], w)
            assert not pooled[pos + len(w):pos + len(w) + 2400].any()
            pos += len(w) + 2400
        assert pos == len(pooled) and np.isfinite(pooled).all()
    finally:
        inference.unload()
