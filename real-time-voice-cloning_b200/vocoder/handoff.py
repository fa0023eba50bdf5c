"""The steps either side of the vocoder call as the reference's callers do them: the
synthesizer hands over one spectrogram per text, the caller concatenates them and remembers the break positions
(toolbox/toolbox.py:263-265), vocodes the concatenation in ONE call (:273), cuts the waveform back at the breaks and puts
0.15 s of silence after every piece (:309-314), then peak-normalises (:321).  Index arithmetic only; the vocoder call is
`vocoder.inference.infer_waveform`.
"""
import numpy as np

from ..config.hparams import sp
from . import inference


def concat_specs(specs):
    """toolbox.py:264-265 -> (spec (80, sum T_i), breaks [T_i])."""
    specs = [np.asarray(s) for s in specs]
    breaks = [int(s.shape[1]) for s in specs]
    return np.concatenate(specs, axis=1), breaks


def add_breaks(wav, breaks, break_seconds=0.15, hop_size=None, sample_rate=None):
    """toolbox.py:309-314.  The vocoder returns (T-1)*hop samples, so the last piece is one hop short (slicing past the end is how
    the reference does it)."""
    hop_size = sp.hop_size if hop_size is None else hop_size
    sample_rate = sp.sample_rate if sample_rate is None else sample_rate
    b_ends = np.cumsum(np.array(breaks) * hop_size)
    b_starts = np.concatenate(([0], b_ends[:-1]))
    wavs = [wav[start:end] for start, end in zip(b_starts, b_ends)]
    gaps = [np.zeros(int(break_seconds * sample_rate))] * len(breaks)
    return np.concatenate([i for w, b in zip(wavs, gaps) for i in (w, b)])


def peak_normalize(wav, peak=0.97):
    """toolbox.py:321."""
    return wav / np.abs(wav).max() * peak


def join_with_gaps(wavs, break_seconds=0.15, sample_rate=None):
    """Pieces that were vocoded separately, each followed by the toolbox's gap (toolbox.py:312-314)."""
    sample_rate = sp.sample_rate if sample_rate is None else sample_rate
    gap = np.zeros(int(break_seconds * sample_rate))
    return np.concatenate([i for w in wavs for i in (np.asarray(w, np.float64), gap)])


def vocode_specs(specs, normalize_peak=True, pooled=False, **infer_kwargs):
    """One text -> one spectrogram each; returns the float64 waveform the toolbox would play.
    pooled=False: the reference's way -- concatenate, vocode once, cut at the breaks (the recurrent state runs across the texts and
    the last piece is one hop short).  pooled=True: every text is its own utterance in ONE engine call (`infer_waveforms`: all
    folds of all texts share the persistent-loop launches, BASELINE config 5), so no text hears its neighbour; piece i has
    (T_i - 1) * hop samples."""
    if pooled:
        wav = join_with_gaps(inference.infer_waveforms(list(specs), **infer_kwargs))
    else:
        spec, breaks = concat_specs(specs)
        wav = add_breaks(inference.infer_waveform(spec, **infer_kwargs), breaks)
    return peak_normalize(wav) if normalize_peak else wav
