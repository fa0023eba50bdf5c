"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, and the
host-only entry points agree with the oracle.  No CUDA calls."""
import os
import re

import numpy as np

from oracle import wavernn_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import __graft_entry__ as g
    g.build()
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    return _native, _native.load()


def test_library_exports_every_declared_symbol():
    _native, lib = _lib()
    header = open(os.path.join(ROOT, "include", "wavernn_b200.h")).read()
    declared = set(re.findall(r"\b(wrnn_[a-z0-9_]+)\s*\(", header)) - {"wrnn_progress_fn"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_native.EXPORTS)


def test_library_is_built_from_these_sources():
    """The shipped .so carries a stamp (build_stamp.json) of the sources and flags it was compiled from; a library that is older than
    its sources must not pass for the product."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("_rtvc_build", os.path.join(ROOT, "real-time-voice-cloning_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()                                      # no-op when the stamp matches
    stamp = json.load(open(b.STAMP))
    assert stamp["sources_sha256"] == b.source_digest()
    assert "arch=compute_100a,code=sm_100a" in stamp["flags"] and "-lineinfo" in stamp["flags"]
    assert os.path.getmtime(b.LIB) >= os.path.getmtime(b.STAMP) - 60


def test_fold_plan_matches_oracle():
    _native, _ = _lib()
    rng = np.random.default_rng(0)
    cases = [(160000, 8000, 800), (960000, 6000, 1000), (960000, 3000, 1500), (4800, 1000, 200), (1000, 1000, 200)]
    cases += [(int(rng.integers(300, 200000)), int(rng.integers(1, 9000)), int(rng.integers(1, 2000))) for _ in range(300)]
    for N, tg, ov in cases:
        if N < ov:
            continue
        assert _native.fold_plan(N, tg, ov) == orc.fold_plan(N, tg, ov), (N, tg, ov)


def test_drop_in_surface():
    import inspect
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import inference
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.vocoder.models.fatchord_version import WaveRNN
    sig = inspect.signature(inference.infer_waveform)
    assert list(sig.parameters)[:6] == ["mel", "normalize", "batched", "target", "overlap", "progress_callback"]
    assert [p.default for p in sig.parameters.values()][:6] == [inspect._empty, True, True, None, None, None]
    gen = inspect.signature(WaveRNN.generate).parameters
    assert list(gen)[1:8] == ["mels", "batched", "target", "overlap", "mu_law", "apply_preemphasis", "progress_callback"]
    assert all(p.default is not inspect._empty for p in list(gen.values())[7:])      # extensions (seed) are optional keywords
    assert list(inspect.signature(inference.load_model).parameters)[:3] == ["weights_fpath", "voc_type", "verbose"]
    assert (base.VOC_TYPE_CPP, base.VOC_TYPE_PYTORCH, base.MODEL_TYPE_FATCHORD) == ("libwavernn", "pytorch", "fatchord-wavernn")
    assert not inference.is_loaded()
    try:
        inference.infer_waveform(np.zeros((80, 30), np.float32))
        raise AssertionError("expected the not-loaded exception")
    except Exception as e:
        assert str(e) == "Please load Wave-RNN in memory before using it"
    for bad in ("nope", "wavernn"):
        try:
            base.init_voc_model(bad, 0)
            raise AssertionError
        except NotImplementedError:
            pass
    # all three model types of the reference (base.py:13-15) have a class; the topologies reject other sizes before touching a GPU
    from rtvc_b200.vocoder.models import geneing_version, runtimeracer_version
    from rtvc_b200.config import hparams
    assert (base.MODEL_TYPE_GENEING, base.MODEL_TYPE_RUNTIMERACER) == ("geneing-wavernn", "runtimeracer-wavernn")
    for cls, hp in ((runtimeracer_version.WaveRNN, hparams.wavernn_fatchord), (geneing_version.WaveRNN, hparams.wavernn_runtimeracer)):
        try:
            cls(rnn_dims=hp.rnn_dims, fc_dims=hp.fc_dims, bits=9, pad=2, upsample_factors=hp.upsample_factors, feat_dims=80,
                compute_dims=hp.compute_dims, res_out_dims=hp.res_out_dims, res_blocks=hp.res_blocks, hop_length=200, sample_rate=16000,
                mode="MOL")
            raise AssertionError
        except NotImplementedError:
            pass
    assert (hparams.wavernn_runtimeracer.gen_target, hparams.wavernn_runtimeracer.gen_overlap) == (6000, 1000)      # config/hparams.py:419-420
    assert (hparams.wavernn_geneing.gen_target, hparams.wavernn_geneing.gen_overlap, hparams.wavernn_geneing.mu_law) == (3000, 1500, False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "real-time-voice-cloning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_auto_precision_rule():
    """The facade's default (PREC_AUTO) picks the loop from the checkpoint and the number of folds in the call."""
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    from rtvc_b200.vocoder.models.fatchord_version import resolve_precision, AUTO_F16_MIN_FOLDS
    assert resolve_precision(512, 0.0, 1) == _native.PREC_F32                      # unbatched: the fp32 loop
    assert resolve_precision(512, 0.0, 5) == _native.PREC_F32                      # a handful of folds: the fp32 loop is as fast and exact
    assert resolve_precision(512, 0.0, 19) == _native.PREC_F16                     # BASELINE config 1: loop_rs.cu (16.1 vs 26.2 us per step)
    assert resolve_precision(512, 0.0, AUTO_F16_MIN_FOLDS) == _native.PREC_F16
    assert resolve_precision(30, 0.0, 213) == _native.PREC_F16                     # infer_waveform(mel) on a 60 s utterance
    assert resolve_precision(1024, 0.0, 1024) == _native.PREC_F16
    assert resolve_precision(256, 0.0, 500) == _native.PREC_F32                    # 8-bit: no tensor-core loop
    assert resolve_precision(512, 0.9, 500) == _native.PREC_SPARSE_F32            # pruned checkpoint (vocoder/pruner.py)
    assert resolve_precision(512, 0.6, 500) == _native.PREC_F16


def test_spectrogram_handoff_matches_the_callers_arithmetic():
    """vocoder/handoff.py against a literal restatement of toolbox/toolbox.py:263-265, 309-314, 321."""
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import handoff
    rng = np.random.default_rng(4)
    specs = [rng.standard_normal((80, t)).astype(np.float32) for t in (37, 5, 112)]
    spec, breaks = handoff.concat_specs(specs)
    assert spec.shape == (80, 154) and breaks == [37, 5, 112]
    wav = rng.standard_normal((154 - 1) * 200)                      # what infer_waveform returns for T = 154
    got = handoff.add_breaks(wav, breaks)
    b_ends = np.cumsum(np.array(breaks) * 200)
    b_starts = np.concatenate(([0], b_ends[:-1]))
    wavs = [wav[s:e] for s, e in zip(b_starts, b_ends)]
    gaps = [np.zeros(int(0.15 * 16000))] * 3
    want = np.concatenate([i for w, b in zip(wavs, gaps) for i in (w, b)])
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert got.shape == ((154 - 1) * 200 + 3 * 2400,)
    n = handoff.peak_normalize(got)
    assert np.array_equal(n, want / np.abs(want).max() * 0.97)
