"""Pins the numpy oracle (oracle/wavernn_oracle.py) against fixtures minted from the UNMODIFIED
reference by oracle/make_golden.py.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import philox, weights, wavernn_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32_10(*c, *k)
        assert tuple(int(g) for g in got) == want


def test_conditioning_matches_reference(golden_dir):
    g = _load(golden_dir, "cond_raw9.npz")
    sd = weights.make_state_dict(seed=int(g["w_seed"]), bits=9, mode="RAW")
    mel = weights.synthetic_mel(int(g["mel_T"]), seed=int(g["mel_seed"])) / np.float32(4.0)
    mels, aux = orc.upsample_network(mel, sd)
    np.testing.assert_allclose(aux[::200], g["aux_frames"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(mels[::37], g["mels_sub"], rtol=0, atol=2e-6)
    # aux is a pure per-frame repeat (Stretch2d): bit-exact structure
    assert np.array_equal(aux, np.repeat(aux[::200], 200, axis=0))


@pytest.mark.parametrize("name,batched", [("gen_raw9_batched.npz", True), ("gen_raw9_unbatched.npz", False)])
def test_generate_raw_index_teacher_forced(golden_dir, name, batched):
    g = _load(golden_dir, name)
    sd = weights.make_state_dict(seed=int(g["w_seed"]), bits=9, mode="RAW")
    mel = weights.synthetic_mel(int(g["mel_T"]), seed=int(g["mel_seed"])) / np.float32(4.0)
    ref_idx = g["index"].astype(np.int64)                       # (B, S-1)
    B, Sm1 = ref_idx.shape
    forced = np.zeros((B, Sm1 + 1), np.float32)
    forced[:, :-1] = orc.label_to_float(ref_idx, 512)
    tg, ov = (int(g["target"]), int(g["overlap"])) if batched else (1000, 200)
    _, tr = orc.generate(mel, sd, mode="RAW", batched=batched, target=tg, overlap=ov, seed=int(g["seed"]),
                         forced_samples=forced, return_trace=True, max_steps=Sm1 + 1)
    agree = (tr["index"][:, :-1] == ref_idx).mean()
    assert agree >= 0.999, agree
    steps = g["logit_steps"]
    np.testing.assert_allclose(tr["logits"][:, steps], g["logits"], rtol=0, atol=2e-5)


def test_generate_raw_free_running_wav(golden_dir):
    """Free-running oracle vs reference: identical indices unless a draw lands within rounding of a
    CDF edge; the final float64 wav must then agree to 1e-9."""
    g = _load(golden_dir, "gen_raw9_batched.npz")
    sd = weights.make_state_dict(seed=int(g["w_seed"]), bits=9, mode="RAW")
    mel = weights.synthetic_mel(int(g["mel_T"]), seed=int(g["mel_seed"])) / np.float32(4.0)
    wav, tr = orc.generate(mel, sd, mode="RAW", batched=True, target=int(g["target"]), overlap=int(g["overlap"]),
                           seed=int(g["seed"]), return_trace=True)
    agree = (tr["index"][:, :-1] == g["index"]).mean()
    assert agree >= 0.999, agree
    assert wav.shape == g["wav"].shape and wav.dtype == np.float64
    if agree == 1.0:
        np.testing.assert_allclose(wav, g["wav"], rtol=0, atol=1e-9)


def test_generate_mol_teacher_forced(golden_dir):
    g = _load(golden_dir, "gen_mol_batched.npz")
    sd = weights.make_state_dict(seed=int(g["w_seed"]), bits=9, mode="MOL")
    mel = weights.synthetic_mel(int(g["mel_T"]), seed=int(g["mel_seed"])) / np.float32(4.0)
    ref = g["samples"]
    B, Sm1 = ref.shape
    forced = np.zeros((B, Sm1 + 1), np.float32)
    forced[:, :-1] = ref
    _, tr = orc.generate(mel, sd, mode="MOL", batched=True, target=int(g["target"]), overlap=int(g["overlap"]),
                         seed=int(g["seed"]), forced_samples=forced, return_trace=True, max_steps=Sm1 + 1)
    assert (tr["index"][:, :-1] == g["mix"][:, :-1]).mean() >= 0.999
    np.testing.assert_allclose(tr["logits"][:, ::8], g["logits_sub"], rtol=0, atol=3e-5)
    same = tr["index"][:, :-1] == g["mix"][:, :-1]
    assert np.abs(tr["samples"][:, :-1] - ref)[same].max() < 1e-4


def test_teacher_forced_forward(golden_dir):
    g = _load(golden_dir, "teacher_forced_raw9.npz")
    sd = weights.make_state_dict(seed=int(g["w_seed"]), bits=9, mode="RAW")
    mel = weights.synthetic_mel(int(g["mel_T"]), seed=int(g["mel_seed"])) / np.float32(4.0)
    xs = np.random.default_rng(int(g["x_seed"])).uniform(-1, 1, size=(1, 22 * 200)).astype(np.float32)[0]
    steps = g["logit_steps"]
    n = int(steps.max()) + 1
    # run only the prefix that is needed up to 1003, then check the tail separately would cost 4400
    # steps of B=1 numpy (~3 s): acceptable.
    out = orc.teacher_forced_logits(xs, mel, sd)
    np.testing.assert_allclose(out[steps], g["logits"], rtol=1e-3, atol=2e-5)
    assert n <= out.shape[0]


def test_fold_unfold_bit_exact(golden_dir):
    g = _load(golden_dir, "fold_unfold.npz")
    for n in range(int(g["n_cases"])):
        N, tg, ov, F, S = (int(v) for v in g["case%d" % n])
        ramp = np.arange(N * 2, dtype=np.float32).reshape(N, 2)
        f = orc.fold_with_overlap(ramp, tg, ov)
        assert f.shape == (F, S, 2)
        assert orc.fold_plan(N, tg, ov)[0] == F
        assert np.array_equal(f[:, 0, 0].astype(np.int64), g["fold_first%d" % n])
        assert np.array_equal(f[:, -1, 1].astype(np.int64), g["fold_last%d" % n])
        assert hashlib.sha256(f.tobytes()).digest() == g["fold_sha%d" % n].tobytes()
        y = np.random.default_rng(100 + n).uniform(-1, 1, size=(F, S))
        un = orc.xfade_and_unfold(y, ov)
        assert hashlib.sha256(un.tobytes()).digest() == g["unfold_sha%d" % n].tobytes()
        if "unfold%d" % n in g:
            assert np.array_equal(un, g["unfold%d" % n])


def test_post_chain(golden_dir):
    g = _load(golden_dir, "post_chain.npz")
    y = g["y"]
    assert np.array_equal(orc.decode_mu_law(y, 512), g["mu512"])
    assert np.array_equal(orc.decode_mu_law(y, 1024), g["mu1024"])
    np.testing.assert_allclose(orc.de_emphasis(y), g["deemph"], rtol=0, atol=1e-12)
    # audio.label_2_float is float64; the path's own float32 form (fatchord_version.py:228) is
    # pinned by the generate tests above.  The two agree to float32 rounding.
    np.testing.assert_allclose(orc.label_to_float(np.arange(512), 512), g["labels"], rtol=0, atol=1e-7)


def test_prune_and_compress(golden_dir):
    g = _load(golden_dir, "prune_compress.npz")
    W, W3 = g["W"], g["W3"]
    sd = {"fc1.weight": W}
    # Linear: one block over all rows (pruner.py:44,62-65)
    S = np.abs(W).reshape(24, 8, 4).sum(2)
    thr = np.sort(S.reshape(-1))[int(24 * 32 // 4 * 0.9)]
    mask = np.repeat((S >= thr).astype(np.float32), 4, axis=1)
    assert np.array_equal(mask, g["mask_lin"])
    # GRU: three gate blocks
    blocks = []
    for b in np.split(W3, 3, axis=0):
        Sb = np.abs(b).reshape(12, 4, 4).sum(2)
        t = np.sort(Sb.reshape(-1))[int(12 * 16 // 4 * 0.75)]
        blocks.append(np.repeat((Sb >= t).astype(np.float32), 4, axis=1))
    assert np.array_equal(np.concatenate(blocks), g["mask_gru"])
    w, idx = orc.compress(W * g["mask_lin"])
    assert np.array_equal(w, g["comp_w"]) and np.array_equal(idx, g["comp_idx"])
    del sd


def test_torch_port_matches_oracle():
    """The CPU-baseline port (torch CPU ops) produces the oracle's samples."""
    from oracle.torch_port import TorchPort
    sd = weights.make_state_dict(seed=11, bits=9, mode="RAW")
    mel = weights.synthetic_mel(24, seed=5) / np.float32(4.0)
    mels, aux = orc.upsample_network(mel, sd)
    mels, aux = orc.fold_with_overlap(mels, 1000, 200), orc.fold_with_overlap(aux, 1000, 200)
    s, done, _ = TorchPort(sd, "RAW").loop(mels, aux, seed=3, max_steps=60)
    _, tr = orc.generate(mel, sd, mode="RAW", batched=True, target=1000, overlap=200, seed=3, return_trace=True,
                         max_steps=60)
    assert done == 60 and (s == tr["samples"]).mean() >= 0.99
    sdm = weights.make_state_dict(seed=12, bits=9, mode="MOL")
    mels, aux = orc.upsample_network(mel, sdm)
    mels, aux = orc.fold_with_overlap(mels, 1000, 200), orc.fold_with_overlap(aux, 1000, 200)
    s, _, _ = TorchPort(sdm, "MOL").loop(mels, aux, seed=3, max_steps=40)
    _, tr = orc.generate(mel, sdm, mode="MOL", batched=True, target=1000, overlap=200, seed=3, return_trace=True,
                         max_steps=40)
    assert (np.abs(s - tr["samples"]) < 1e-3).mean() >= 0.9      # free-running: a mixture flip forks the trajectory
