"""Parity of the CUDA path (through the C ABI) against the reference's golden vectors and the oracle.
Tolerances (BASELINE.json north_star): teacher-forced logits within 1e-3 relative (fp32); identical
sample indices on >= 99.9 % of steps under the same noise; fold/unfold indices bit-exact."""
import hashlib

import numpy as np
import pytest

from oracle import wavernn_oracle as orc
from tests.util import golden, make_model, norm_mel

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3
AGREE = 0.999


@pytest.fixture(scope="module")
def raw9():
    return make_model(seed=11, bits=9, mode="RAW")


@pytest.fixture(scope="module")
def mol():
    return make_model(seed=12, bits=9, mode="MOL")


def rel_err(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_conditioning_matches_reference(raw9):
    model, sd = raw9
    g = golden("cond_raw9.npz")
    mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
    up, aux = model.conditioning(mel)
    np.testing.assert_allclose(aux, g["aux_frames"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(up[::37], g["mels_sub"], rtol=0, atol=2e-6)


def test_xfade_unfold_bit_exact(raw9):
    model, _ = raw9
    g = golden("fold_unfold.npz")
    for n in range(int(g["n_cases"])):
        N, tg, ov, F, S = (int(v) for v in g["case%d" % n])
        y = np.random.default_rng(100 + n).uniform(-1, 1, size=(F, S))
        un = model.xfade_and_unfold(y, tg, ov)
        assert hashlib.sha256(un.tobytes()).digest() == g["unfold_sha%d" % n].tobytes(), (N, tg, ov)


def test_fold_with_overlap_bit_exact(raw9):
    import torch
    model, _ = raw9
    g = golden("fold_unfold.npz")
    for n in range(int(g["n_cases"])):
        N, tg, ov, F, S = (int(v) for v in g["case%d" % n])
        ramp = torch.arange(N * 2, dtype=torch.float32).reshape(1, N, 2)
        f = model.fold_with_overlap(ramp, tg, ov).numpy()
        assert f.shape == (F, S, 2)
        assert hashlib.sha256(f.tobytes()).digest() == g["fold_sha%d" % n].tobytes(), (N, tg, ov)


@pytest.mark.parametrize("name,batched", [("gen_raw9_batched.npz", True), ("gen_raw9_unbatched.npz", False)])
def test_raw_teacher_forced_vs_reference(raw9, name, batched):
    """Feed the kernel the REFERENCE's samples; compare its logits and its own draws step by step."""
    model, _ = raw9
    g = golden(name)
    mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
    ref_idx = g["index"].astype(np.int64)
    B, Sm1 = ref_idx.shape
    forced = np.zeros((B, Sm1 + 1), np.float32)
    forced[:, :-1] = orc.label_to_float(ref_idx, 512)
    tg, ov = (int(g["target"]), int(g["overlap"])) if batched else (0, 0)
    out = model.generate_debug(mel, batched, tg, ov, forced=forced, want_logits=True, seed=int(g["seed"]))
    steps = g["logit_steps"]
    assert rel_err(out["logits"][:, steps], g["logits"]) < REL_TOL
    mine = np.rint((out["samples"] + 1.0) * 511 / 2.0).astype(np.int64)
    agree = float((mine[:, :-1] == ref_idx).mean())
    assert agree >= AGREE, agree


def test_raw_free_running_vs_reference(raw9):
    model, _ = raw9
    g = golden("gen_raw9_batched.npz")
    mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
    model.seed = int(g["seed"])
    wav = model.generate(mel[None], True, int(g["target"]), int(g["overlap"]), True, True)
    out = model.generate_debug(mel, True, int(g["target"]), int(g["overlap"]), seed=int(g["seed"]))
    mine = np.rint((out["samples"] + 1.0) * 511 / 2.0).astype(np.int64)
    agree = float((mine[:, :-1] == g["index"]).mean())
    assert agree >= AGREE, agree
    assert wav.dtype == np.float64 and wav.shape == g["wav"].shape
    if agree == 1.0:
        np.testing.assert_allclose(wav, g["wav"], rtol=0, atol=1e-9)


def test_mol_teacher_forced_vs_reference(mol):
    model, _ = mol
    g = golden("gen_mol_batched.npz")
    mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
    ref = g["samples"]
    B, Sm1 = ref.shape
    forced = np.zeros((B, Sm1 + 1), np.float32)
    forced[:, :-1] = ref
    out = model.generate_debug(mel, True, int(g["target"]), int(g["overlap"]), forced=forced, want_logits=True,
                               seed=int(g["seed"]))
    assert rel_err(out["logits"][:, ::8], g["logits_sub"]) < REL_TOL
    d = np.abs(out["samples"][:, :-1] - ref)
    assert float((d < 1e-4).mean()) >= AGREE


def test_mol_free_running_wav(mol):
    model, _ = mol
    g = golden("gen_mol_batched.npz")
    mel = norm_mel(int(g["mel_T"]), int(g["mel_seed"]))
    model.seed = int(g["seed"])
    wav = model.generate(mel[None], True, int(g["target"]), int(g["overlap"]), True, True)
    assert wav.shape == g["wav"].shape and np.isfinite(wav).all()
    out = model.generate_debug(mel, True, int(g["target"]), int(g["overlap"]), seed=int(g["seed"]))
    d = np.abs(out["samples"][:, :-1] - g["samples"])
    if float(d.max()) < 1e-4:           # no mixture flip anywhere: the float64 wav must match too
        np.testing.assert_allclose(wav, g["wav"], rtol=0, atol=1e-3)


def test_raw_vs_oracle_other_seed(raw9):
    """Fresh input not in the fixtures: oracle teacher-forced on the kernel's samples."""
    model, sd = raw9
    mel = norm_mel(23, 77)
    out = model.generate_debug(mel, True, 600, 100, want_logits=True, seed=1234, max_steps=300)
    _, tr = orc.generate(mel, sd, mode="RAW", batched=True, target=600, overlap=100, seed=1234,
                         forced_samples=np.pad(out["samples"], ((0, 0), (0, 500))), return_trace=True, max_steps=300)
    assert rel_err(out["logits"], tr["logits"]) < REL_TOL
    mine = np.rint((out["samples"] + 1.0) * 511 / 2.0).astype(np.int64)
    assert float((mine == tr["index"]).mean()) >= AGREE


def test_post_chain_vs_oracle(raw9):
    model, _ = raw9
    rng = np.random.default_rng(5)
    F, tg, ov, T = 5, 1000, 200, 30
    S = tg + 2 * ov
    k = rng.integers(0, 512, size=(F, S))
    samples = orc.label_to_float(k, 512)
    for mu_law in (True, False):
        for pre in (True, False):
            wav = model.postprocess(samples, True, ov, T, mu_law, pre)
            want = orc.finish(orc.xfade_and_unfold(samples.astype(np.float64), ov), (T - 1) * 200, 512, mu_law, pre)
            np.testing.assert_allclose(wav, want, rtol=0, atol=1e-11)
    # unbatched
    s1 = orc.label_to_float(rng.integers(0, 512, size=(1, T * 200)), 512)
    wav = model.postprocess(s1, False, 0, T, True, True)
    np.testing.assert_allclose(wav, orc.finish(s1[0].astype(np.float64), (T - 1) * 200, 512, True, True), rtol=0, atol=1e-11)


def test_fold_sharding_is_invisible(raw9):
    """Multi-GPU partition property: running fold ranges separately gives bit-identical samples."""
    model, _ = raw9
    mel = norm_mel(40, 3)
    full = model.generate_debug(mel, True, 500, 100, seed=9, max_steps=200)["samples"]
    a = model.generate_debug(mel, True, 500, 100, seed=9, max_steps=200, fold_begin=0, fold_end=5)["samples"]
    b = model.generate_debug(mel, True, 500, 100, seed=9, max_steps=200, fold_begin=5, fold_end=99)["samples"]
    assert np.array_equal(np.concatenate([a, b]), full)


def test_batch_of_utterances_matches_single(raw9):
    model, _ = raw9
    mels = [norm_mel(22, 1), norm_mel(31, 2), norm_mel(26, 3)]
    model.seed = 5
    wavs = model.generate_batch(mels, True, 800, 100, True, True)
    for i, m in enumerate(mels):
        rq, arrs, wav, offsets, keep = model._request([m], True, 800, 100, True, True, None, utt_index0=i)
        model._run(rq)
        assert np.array_equal(wav[:len(wavs[i])], wavs[i])


def test_errors(raw9):
    model, _ = raw9
    with pytest.raises(ValueError):
        model.generate(norm_mel(20, 1)[None], True, 1000, 200, True, True)      # Q8: T <= 20
    with pytest.raises(ValueError):
        model.generate(norm_mel(30, 1)[None], True, 1000, 0, True, True)        # overlap == 0
    import rtvc_b200.vocoder.inference as inf
    inf.unload()
    with pytest.raises(Exception, match="Please load Wave-RNN in memory before using it"):
        inf.infer_waveform(np.zeros((80, 30), np.float32))


def test_full_size_config1_properties(raw9):
    """BASELINE config 1 at full size (19 folds x 9600 steps): shape, determinism, range."""
    model, _ = raw9
    mel = norm_mel(800, 1)
    model.seed = 1
    w1 = model.generate(mel[None], True, 8000, 800, True, True)
    t = dict(model.last_timings)
    assert t["n_folds"] == 19 and t["n_steps"] == 9600
    w2 = model.generate(mel[None], True, 8000, 800, True, True)
    assert w1.shape == (159800,) and w1.dtype == np.float64
    assert np.array_equal(w1, w2)
    assert np.isfinite(w1).all() and w1[-1] == 0.0 and np.all(w1[:400] == 0.0)   # Q7: first overlap//2 muted


def test_raw_10bit_and_8bit_vs_oracle():
    """bits = 10 is the reference's default (config/hparams.py:223); 8 exercises the smallest class count."""
    for bits, C in ((10, 1024), (8, 256)):
        model, sd = make_model(seed=21, bits=bits, mode="RAW")
        mel = norm_mel(23, 8)
        out = model.generate_debug(mel, True, 600, 100, want_logits=True, seed=77, max_steps=150)
        forced = np.pad(out["samples"], ((0, 0), (0, 800 - 150)))
        _, tr = orc.generate(mel, sd, mode="RAW", batched=True, target=600, overlap=100, seed=77, forced_samples=forced,
                             return_trace=True, max_steps=150)
        assert rel_err(out["logits"], tr["logits"]) < REL_TOL
        mine = np.rint((out["samples"] + 1.0) * (C - 1) / 2.0).astype(np.int64)
        assert float((mine == tr["index"]).mean()) >= AGREE


def test_mol_unbatched_vs_oracle(mol):
    model, sd = mol
    mel = norm_mel(22, 6)
    out = model.generate_debug(mel, False, 0, 0, want_logits=True, seed=5, max_steps=300)
    forced = np.pad(out["samples"], ((0, 0), (0, 22 * 200 - 300)))
    _, tr = orc.generate(mel, sd, mode="MOL", batched=False, seed=5, forced_samples=forced, return_trace=True, max_steps=300)
    assert rel_err(out["logits"], tr["logits"]) < REL_TOL
    assert float((np.abs(out["samples"] - tr["samples"]) < 1e-4).mean()) >= AGREE


def test_ragged_batch_all_precisions(raw9):
    """Utterances of different lengths pooled in one call: every loop gives every utterance its own length and noise."""
    model, _ = raw9
    mels = [norm_mel(T, 40 + T) for T in (21, 57, 33, 90)]
    model.seed = 9
    ref = model.generate_batch(mels, True, 700, 150, True, True)
    assert [len(w) for w in ref] == [(m.shape[1] - 1) * 200 for m in mels]
    model.precision = 1
    f16 = model.generate_batch(mels, True, 700, 150, True, True)
    model.precision = 0
    for a, b in zip(ref, f16):
        assert a.shape == b.shape and np.isfinite(b).all()
        assert np.mean(np.abs(a - b) < 1e-6) > 0.5       # same draws until the first fp16-induced flip in a fold


def test_progress_callback_contract(raw9):
    model, _ = raw9
    calls = []
    wav = model.generate(norm_mel(30, 2)[None], True, 1000, 200, True, True,
                         progress_callback=lambda i, seq_len, b_size, rate: calls.append((i, seq_len, b_size, rate)))
    assert wav.shape == (29 * 200,)
    assert calls and calls[-1][0] == 1399 and calls[-1][1] == 1400 and calls[-1][2] == 5 and calls[-1][3] > 0


def test_facade_default_precision_is_auto(mol):
    """load_model / load_state leave the engine on PREC_AUTO: the tensor-core loop for calls with many folds, the fp32 loop for
    few folds and for unbatched generation (fatchord_version.resolve_precision)."""
    import copy
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    from rtvc_b200.config.hparams import wavernn_fatchord
    from rtvc_b200.vocoder import inference
    from tests.util import norm_mel
    _, sd = mol
    hp = copy.deepcopy(wavernn_fatchord)
    hp.bits, hp.mode = 9, "MOL"
    m = inference.load_state(sd, override_hp_fatchord=hp)
    assert m.precision == _native.PREC_AUTO
    w = inference.infer_waveform(norm_mel(960, 1) * 4.0)                   # 12 s, the facade's own fold plan (3000 / 1500): 42 folds
    assert w.shape == ((960 - 1) * 200,) and np.isfinite(w).all()
    assert m.last_timings["n_folds"] >= 24 and m.last_timings["precision"] == _native.PREC_F16
    inference.infer_waveform(norm_mel(60, 1) * 4.0)                        # 0.75 s: 3 folds
    assert m.last_timings["n_folds"] < 4 and m.last_timings["precision"] == _native.PREC_F32
    inference.infer_waveform(norm_mel(40, 1) * 4.0, batched=False)
    assert m.last_timings["precision"] == _native.PREC_F32
