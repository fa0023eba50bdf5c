"""The role-specialised tensor-core loop (csrc/loop_rs.cu; MOL, <= 128 folds per 48-CTA group) against the ORACLE at the
full sizes BASELINE config 3 is run at, and the full-size tensor-core paths of loop_tc.cu (CTA pairs, sampler CTAs, four
pipelined sets) against the oracle as well.  Gates (BASELINE.json north_star): teacher-forced logits within 1e-3 relative,
>= 99.9 % of the draws identical (MOL: the continuous sample within 1e-3 of the oracle's under the same noise).
Every test prints the agreement it achieved."""
import numpy as np
import pytest

from oracle import wavernn_oracle as orc
from tests.util import make_model, norm_mel

pytestmark = pytest.mark.gpu
F16 = 1
REL_TOL = 1e-3
AGREE = 0.999


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _oracle_check(model, sd, mode, mel, target, overlap, steps, seed, tag):
    """Kernel free-running for `steps` steps; the oracle teacher-forced on the KERNEL's samples: per-step logits and draws."""
    out = model.generate_debug(mel, True, target, overlap, want_logits=True, seed=seed, max_steps=steps, precision=F16)
    F = out["samples"].shape[0]
    S = target + 2 * overlap
    forced = np.zeros((F, S), np.float32)
    forced[:, :steps] = out["samples"]
    _, tr = orc.generate(mel, sd, mode=mode, batched=True, target=target, overlap=overlap, seed=seed, forced_samples=forced,
                         return_trace=True, max_steps=steps)
    err = _rel(out["logits"], tr["logits"])
    if mode == "RAW":
        C = sd["fc3.weight"].shape[0]
        mine = np.rint((out["samples"] + 1.0) * (C - 1) / 2.0).astype(np.int64)
        agree = float((mine == tr["index"]).mean())
    else:
        agree = float((np.abs(out["samples"] - tr["samples"]) < 1e-3).mean())
    print("%s: %d folds x %d steps vs oracle: logits rel err %.3e, draw agreement %.5f" % (tag, F, steps, err, agree))
    assert err < REL_TOL, err
    assert agree >= AGREE, agree
    return out


@pytest.mark.parametrize("target,overlap,folds", [(3000, 1500, 213), (6000, 1000, 137)])
def test_rs_loop_config3_full_size_vs_oracle(target, overlap, folds):
    """BASELINE config 3 (MOL, 60 s) on the two fold plans SURVEY.md 8(d) names: infer_waveform's own default (3000 / 1500)
    and 6000 / 1000; first 80 steps of all folds (two groups, in-kernel expanders, conditioning ring)."""
    model, sd = make_model(seed=12, bits=9, mode="MOL")
    mel = norm_mel(4800, 1)
    out = _oracle_check(model, sd, "MOL", mel, target, overlap, 80, 9, "loop_rs cfg3 %d/%d" % (target, overlap))
    assert out["samples"].shape[0] == folds
    assert dict(model.last_timings)["n_folds"] == folds


@pytest.mark.parametrize("mode,seed,bits,T,target,overlap,folds", [("MOL", 12, 9, 4800, 3000, 1500, 213), ("RAW", 11, 9, 800, 8000, 800, 19)])
def test_rs_loop_full_length_vs_fp32_loop(mode, seed, bits, T, target, overlap, folds):
    """EVERY step of the two headline calls (config 3 on infer_waveform's own plan: 213 folds x 6000 steps; config 1: 19 folds x 9600
    steps) -- all phases of all frames, the fold tails past the utterance, 30 / 48 generations of the exchange buffers' bit: the fp16
    loop free-running, the oracle-anchored fp32 loop teacher-forced on its samples; logits within 1e-3 of the largest logit, draws >= 99.9 %."""
    model, _ = make_model(seed=seed, bits=bits, mode=mode)
    mel = norm_mel(T, 1)
    S = target + 2 * overlap
    b = model.generate_debug(mel, True, target, overlap, want_logits=True, seed=9, precision=F16)
    assert b["samples"].shape == (folds, S) and dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rs_kernel"
    a = model.generate_debug(mel, True, target, overlap, forced=b["samples"], want_logits=True, seed=9, precision=0)
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_f32_kernel"
    err = _rel(b["logits"], a["logits"])
    agree = float((a["samples"] == b["samples"]).mean()) if mode == "RAW" else float((np.abs(a["samples"] - b["samples"]) < 1e-3).mean())
    worst = float(np.abs(b["logits"] - a["logits"]).max(axis=(0, 2)).argmax())
    print("loop_rs %s full length, %d folds x %d steps vs fp32 loop: logits rel err %.3e (worst step %d), draw agreement %.5f" % (mode, folds, S, err, worst, agree))
    assert np.isfinite(b["logits"]).all()
    assert err < REL_TOL and agree >= AGREE, (err, agree)


def test_rs_loop_three_full_groups_vs_oracle():
    """257 .. 384 folds: three groups of up to 128 folds each (only possible since no SM is needed for expanders) against the ORACLE."""
    model, sd = make_model(seed=12, bits=9, mode="MOL")
    mel = norm_mel(4800, 2)
    out = _oracle_check(model, sd, "MOL", mel, 2200, 360, 48, 9, "loop_rs three groups 2200/360")
    assert 257 <= out["samples"].shape[0] <= 384, out["samples"].shape
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rs_kernel"


def test_rs_loop_two_waves_vs_oracle(monkeypatch):
    """385 .. 768 MOL folds run as two balanced waves of the role-specialised loop (measured faster than one loop_tc launch):
    against the ORACLE, and WRNN_RS_WAVES=0 puts the same call back on loop_tc."""
    model, sd = make_model(seed=12, bits=9, mode="MOL")
    mel = norm_mel(4800, 2)
    out = _oracle_check(model, sd, "MOL", mel, 1705, 170, 32, 9, "loop_rs two waves 1705/170")
    t = dict(model.last_timings)
    assert out["samples"].shape[0] == 512 and t["loop_kernel"] == "wrnn_loop_rs_kernel" and t["n_launches"] == 2, t
    monkeypatch.setenv("WRNN_RS_WAVES", "0")
    model.generate_debug(mel, True, 1705, 170, seed=9, max_steps=8, precision=F16)
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_tc_kernel"


@pytest.mark.parametrize("bits,T,target,overlap,folds,steps", [(9, 800, 8000, 800, 19, 96), (10, 4800, 3000, 1500, 213, 64), (9, 4800, 6000, 1000, 137, 64)])
def test_rs_loop_raw_vs_oracle(bits, T, target, overlap, folds, steps):
    """RAW on the role-specialised loop (sampler CTAs: fc3 slices, soft-max partials exchanged between the CTAs, inverse-CDF draw):
    BASELINE config 1 (9-bit, 10 s, 19 folds) and a 60 s utterance on the two config-3 fold plans (10-bit: the reference's default
    bits, 8 sampler CTAs per group; 9-bit: 4) against the ORACLE teacher-forced on the kernel's samples."""
    model, sd = make_model(seed=11, bits=bits, mode="RAW")
    mel = norm_mel(T, 1)
    out = _oracle_check(model, sd, "RAW", mel, target, overlap, steps, 9, "loop_rs RAW %d-bit %d/%d" % (bits, target, overlap))
    assert out["samples"].shape[0] == folds
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rs_kernel"


def test_rs_loop_raw_deterministic_groups_and_generate(monkeypatch):
    """RAW: same samples run to run and with one / two groups; the public generate() returns the right length."""
    model, _ = make_model(seed=11, bits=9, mode="RAW")
    mel = norm_mel(400, 3)
    a = model.generate_debug(mel, True, 700, 150, seed=4, max_steps=300, precision=F16)
    b = model.generate_debug(mel, True, 700, 150, seed=4, max_steps=300, precision=F16)
    np.testing.assert_array_equal(a["samples"], b["samples"])
    assert a["samples"].shape[0] == 94
    monkeypatch.setenv("WRNN_RS_GROUPS", "1")
    c = model.generate_debug(mel, True, 700, 150, seed=4, max_steps=300, precision=F16)
    np.testing.assert_array_equal(a["samples"], c["samples"])
    monkeypatch.delenv("WRNN_RS_GROUPS")
    model.precision = F16
    model.seed = 3
    w = model.generate(mel[None], True, 3000, 1500, True, True)
    assert w.shape == ((400 - 1) * 200,) and w.dtype == np.float64 and np.isfinite(w).all()
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rs_kernel"


def test_rs_loop_deterministic_and_groups_invisible(monkeypatch):
    """Inline conditioning (the default: per-frame rows + the mel share inside the MMA): same samples run to run and with two and
    three CTA groups, with and without the grid padded to the SM count."""
    model, _ = make_model(seed=12, bits=9, mode="MOL")
    mel = norm_mel(600, 3)                                   # 120000 samples, 700 + 2 x 150 -> 141 folds
    a = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    b = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    assert a["samples"].shape[0] == 141
    np.testing.assert_array_equal(a["samples"], b["samples"])
    for g in (2, 3):
        monkeypatch.setenv("WRNN_RS_GROUPS", str(g))
        e = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
        np.testing.assert_array_equal(a["samples"], e["samples"])
    monkeypatch.setenv("WRNN_RS_PAD", "0")
    f = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    np.testing.assert_array_equal(a["samples"], f["samples"])
    monkeypatch.delenv("WRNN_RS_PAD")


def test_rs_loop_layout_calibration_is_invisible(monkeypatch):
    """The first long call of a shape on an engine times a few layouts (groups x grid padding) and keeps the fastest: the samples
    are the same with the calibration (first call: trials + run; second call: the kept layout) and without it, RAW and MOL."""
    for mode, seed in (("RAW", 11), ("MOL", 12)):
        mel = norm_mel(300, 3)                                   # 60000 samples, 1500 + 2 x 300 -> 34 folds x 2100 steps
        monkeypatch.setenv("WRNN_RS_CALIBRATE", "0")
        ref_model, _ = make_model(seed=seed, bits=9, mode=mode)
        a = ref_model.generate_debug(mel, True, 1500, 300, seed=4, precision=F16)
        monkeypatch.delenv("WRNN_RS_CALIBRATE")
        model, _ = make_model(seed=seed, bits=9, mode=mode)
        b = model.generate_debug(mel, True, 1500, 300, seed=4, precision=F16)
        c = model.generate_debug(mel, True, 1500, 300, seed=4, precision=F16)
        assert a["samples"].shape[1] == 2100 and dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rs_kernel"
        np.testing.assert_array_equal(a["samples"], b["samples"])
        np.testing.assert_array_equal(a["samples"], c["samples"])


def test_rs_loop_record_ring_path(monkeypatch):
    """WRNN_RS_INLINE=0 keeps the round's first form of the loop -- conditioning records from expander CTAs through an L2-resident
    ring -- selectable: against the ORACLE on the 213-fold plan, and the same samples (a) run to run, (b) with the records expanded
    up front, (c) with a ring so small that it wraps every 12 steps, (d) with one and two groups."""
    monkeypatch.setenv("WRNN_RS_INLINE", "0")
    model, sd = make_model(seed=12, bits=9, mode="MOL")
    _oracle_check(model, sd, "MOL", norm_mel(4800, 1), 3000, 1500, 64, 9, "loop_rs (record ring) cfg3 3000/1500")
    mel = norm_mel(500, 3)                                   # 100000 samples, 700 + 2 x 150 -> 118 folds
    a = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    b = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    assert a["samples"].shape[0] == 118
    np.testing.assert_array_equal(a["samples"], b["samples"])
    monkeypatch.setenv("WRNN_RS_EXPAND", "0")
    c = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    np.testing.assert_array_equal(a["samples"], c["samples"])
    monkeypatch.delenv("WRNN_RS_EXPAND")
    monkeypatch.setenv("WRNN_RS_RING_MB", "1")
    d = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
    np.testing.assert_array_equal(a["samples"], d["samples"])
    monkeypatch.delenv("WRNN_RS_RING_MB")
    for g in (1, 2):
        monkeypatch.setenv("WRNN_RS_GROUPS", str(g))
        e = model.generate_debug(mel, True, 700, 150, want_logits=False, seed=4, max_steps=400, precision=F16)
        np.testing.assert_array_equal(a["samples"], e["samples"])


def test_rs_loop_inline_conditioning_tail_and_multi_utterance():
    """Inline conditioning at its edges: folds whose tail runs past the utterance (rows of frame >= T: bias-only row + the zero mel
    row, Q9), a fold that starts exactly on a frame boundary and one that does not, and several utterances in one call (row spaces
    of different utterances) -- free-running fp16 samples teacher-force the fp32 loop: logits within 1e-3, samples within 1e-3 on >= 99.9 %."""
    model, _ = make_model(seed=12, bits=9, mode="MOL")
    for T, tg, ov in [(31, 1000, 200), (57, 830, 170), (140, 2600, 300)]:       # (200 T) is not a multiple of target + overlap: padded tails
        mel = norm_mel(T, 7)
        S = tg + 2 * ov
        b = model.generate_debug(mel, True, tg, ov, want_logits=True, seed=6, precision=F16)
        a = model.generate_debug(mel, True, tg, ov, forced=np.pad(b["samples"], ((0, 0), (0, S - b["samples"].shape[1]))), want_logits=True, seed=6)   # fp32 loop
        err = _rel(b["logits"], a["logits"])
        agree = float((np.abs(a["samples"] - b["samples"]) < 1e-3).mean())
        print("loop_rs inline, %d folds x %d steps (padded tail) vs fp32 loop: logits rel err %.3e, agreement %.5f" % (b["samples"].shape[0], b["samples"].shape[1], err, agree))
        assert dict(model.last_timings)["n_steps"] == S
        assert err < REL_TOL and agree >= AGREE, (T, err, agree)
    # several utterances: each must equal the same utterance vocoded alone (same Philox counters: utterance index is a key)
    model.precision = F16
    mels = [norm_mel(T, 20 + i) for i, T in enumerate((40, 64, 33))]
    wavs = model.generate_batch([m[None] for m in mels], True, 800, 200, True, True, seed=5)
    for i, m in enumerate(mels):
        assert wavs[i].shape == ((m.shape[1] - 1) * 200,) and np.isfinite(wavs[i]).all()
        alone = model.generate_batch([m[None]], True, 800, 200, True, True, seed=5, utt_index0=i)[0]
        np.testing.assert_array_equal(wavs[i], alone)
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_rs_kernel"


def test_rs_loop_small_and_ragged_fold_counts():
    """1, 5, 33 and 129 folds (partial quadrants, a single group, two groups with an odd split), teacher-forced on the fp32
    loop's samples: logits within 1e-3 of the fp32 loop, samples within 1e-3 on >= 99.9 %."""
    model, _ = make_model(seed=12, bits=9, mode="MOL")
    for T, tg, ov in [(6, 800, 200), (26, 800, 200), (166, 800, 200), (646, 800, 200)]:
        mel = norm_mel(T, 5)
        a = model.generate_debug(mel, True, tg, ov, want_logits=True, seed=6, max_steps=64)                       # fp32 loop
        forced = np.pad(a["samples"], ((0, 0), (0, tg + 2 * ov - 64)))
        b = model.generate_debug(mel, True, tg, ov, forced=forced, want_logits=True, seed=6, max_steps=64, precision=F16)
        err = _rel(b["logits"], a["logits"])
        agree = float((np.abs(a["samples"] - b["samples"]) < 1e-3).mean())
        print("loop_rs %d folds vs fp32 loop: logits rel err %.3e, agreement %.5f" % (a["samples"].shape[0], err, agree))
        assert err < REL_TOL and agree >= AGREE, (a["samples"].shape[0], err, agree)


def test_rs_loop_full_generate_matches_loop_tc_shape_and_is_finite(monkeypatch):
    """The public generate() through the role-specialised loop: right length, float64, finite, deterministic; and the same
    call through loop_tc.cu (WRNN_RS=0) gives a waveform that agrees until the first differing draw."""
    model, _ = make_model(seed=12, bits=9, mode="MOL")
    model.precision = F16
    mel = norm_mel(400, 2)
    model.seed = 3
    w1 = model.generate(mel[None], True, 3000, 1500, True, True)
    w2 = model.generate(mel[None], True, 3000, 1500, True, True)
    assert w1.shape == ((400 - 1) * 200,) and w1.dtype == np.float64 and np.isfinite(w1).all()
    assert np.array_equal(w1, w2)


@pytest.mark.parametrize("mode,seed", [("MOL", 12), ("RAW", 11)])
def test_loop_tc_full_size_cfg3_plan_vs_oracle(mode, seed):
    """loop_tc.cu at the size its large-fold paths exist for (1024 folds: four pipelined sets, CTA pairs, sampler CTAs,
    in-kernel expanders) against the ORACLE: first 64 steps of the 853 / 85 plan.  RAW is the kernel configuration a pooled
    cfg5 wave of 1024 folds runs (RAW-512 sampler CTAs as CTA pairs)."""
    model, sd = make_model(seed=seed, bits=9, mode=mode)
    mel = norm_mel(4800, 1)
    out = _oracle_check(model, sd, mode, mel, 853, 85, 64, 9, "loop_tc cfg3 853/85 %s" % mode)
    assert out["samples"].shape[0] == 1024
