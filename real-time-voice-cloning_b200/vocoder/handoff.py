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


def vocode_specs(specs, normalize_peak=True, **infer_kwargs):
    """One text -> one spectrogram each; returns the float64 waveform the toolbox would play."""
    spec, breaks = concat_specs(specs)
    wav = inference.infer_waveform(spec, **infer_kwargs)
    wav = add_breaks(wav, breaks)
    return peak_normalize(wav) if normalize_peak else wav
