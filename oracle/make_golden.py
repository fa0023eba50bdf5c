"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

TEST INFRASTRUCTURE.  `python -m oracle.make_golden` imports /root/reference (oracle/ref_import.py),
loads the deterministic weights of oracle/weights.py into the reference's own WaveRNN, patches ONLY
the two random draws (torch.distributions.Categorical.sample -> inverse-CDF on a Philox uniform;
fatchord_version.sample_from_discretized_mix_logistic -> the same function body as
vocoder/distribution.py:104-140 with its two uniform_() calls replaced by Philox uniforms), runs the
reference's generate()/forward()/fold/unfold/audio helpers and writes small .npz fixtures under
tests/golden/.  /root/reference does not exist on the GPU box, hence the committed fixtures.
"""
import copy
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import philox, weights  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build_reference_model(base, hparams, sd, bits, mode):
    hp = copy.deepcopy(hparams.wavernn_fatchord)
    hp.bits = bits
    hp.mode = mode
    model, _ = base.init_voc_model(base.MODEL_TYPE_FATCHORD, torch.device("cpu"), override_hp_fatchord=hp)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return model.eval()


class NoiseInjector:
    """Replays the build-defined Philox noise inside the reference's generate()."""

    def __init__(self, fv, seed, utt=0):
        self.fv, self.seed, self.utt, self.step = fv, seed, utt, 0
        self.mix_idx = []

    def __enter__(self):
        inj = self
        self._orig_sample = torch.distributions.Categorical.sample
        self._orig_mol = self.fv.sample_from_discretized_mix_logistic

        def cat_sample(dist, sample_shape=torch.Size()):
            probs = dist.probs                                     # (B, C) renormalised by Categorical
            B = probs.shape[0]
            u = philox.raw_uniforms(inj.seed, inj.step + 1, B, utt=inj.utt)[inj.step]
            inj.step += 1
            cdf = torch.cumsum(probs, dim=1)
            k = (cdf < torch.from_numpy(u)[:, None]).sum(dim=1)
            return torch.clamp(k, max=probs.shape[1] - 1)

        def mol_sample(y, log_scale_min=None):
            # body of vocoder/distribution.py:104-140, uniform_() -> injected
            if log_scale_min is None:
                log_scale_min = float(np.log(1e-14))
            nr_mix = y.size(1) // 3
            y = y.transpose(1, 2)
            logit_probs = y[:, :, :nr_mix]
            B = y.shape[1]
            um, ul = philox.mol_uniforms(inj.seed, inj.step + 1, B, utt=inj.utt)
            um, ul = torch.from_numpy(um[inj.step])[None], torch.from_numpy(ul[inj.step])[None]
            inj.step += 1
            temp = logit_probs.data - torch.log(-torch.log(um))
            _, argmax = temp.max(dim=-1)
            inj.mix_idx.append(argmax.view(-1).numpy().copy())
            one_hot = inj.fv.to_one_hot(argmax, nr_mix) if hasattr(inj.fv, "to_one_hot") else \
                torch.nn.functional.one_hot(argmax, nr_mix).float()
            means = torch.sum(y[:, :, nr_mix:2 * nr_mix] * one_hot, dim=-1)
            log_scales = torch.clamp(torch.sum(y[:, :, 2 * nr_mix:3 * nr_mix] * one_hot, dim=-1), min=log_scale_min)
            x = means + torch.exp(log_scales) * (torch.log(ul) - torch.log(1. - ul))
            return torch.clamp(torch.clamp(x, min=-1.), max=1.)

        torch.distributions.Categorical.sample = cat_sample
        self.fv.sample_from_discretized_mix_logistic = mol_sample
        return self

    def __exit__(self, *a):
        torch.distributions.Categorical.sample = self._orig_sample
        self.fv.sample_from_discretized_mix_logistic = self._orig_mol


def run_generate(model, fv, mel_norm, seed, batched, target, overlap, mu_law=True, preemph=True):
    logits, fed = [], []
    h1 = model.fc3.register_forward_hook(lambda m, i, o: logits.append(o.detach().numpy().copy()))
    h2 = model.I.register_forward_hook(lambda m, i, o: fed.append(i[0][:, 0].detach().numpy().copy()))
    try:
        with NoiseInjector(fv, seed) as inj:
            wav = model.generate(torch.from_numpy(mel_norm[None]), batched, target, overlap, mu_law, preemph,
                                 progress_callback=lambda *a: None)
    finally:
        h1.remove()
        h2.remove()
    model.eval()                                   # generate() leaves train mode on (Q1)
    logits = np.stack(logits, axis=1)              # (B, S, C)
    fed = np.stack(fed, axis=1)                    # (B, S) the x fed at each step (x[0] = 0)
    samples = np.concatenate([fed[:, 1:], np.zeros((fed.shape[0], 1), np.float32)], axis=1)
    mix = np.stack(inj.mix_idx, axis=1) if inj.mix_idx else None
    return wav, logits, fed, samples, mix


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    base, fv, hparams, _inference = import_reference()
    from vocoder import audio as ref_audio
    from vocoder.pruner import PruneMask
    from vocoder.libwavernn import convert as ref_convert

    # ---- G1: RAW 9-bit batched ------------------------------------------------------------------
    sd = weights.make_state_dict(seed=11, bits=9, mode="RAW")
    model = build_reference_model(base, hparams, sd, 9, "RAW")
    mel = weights.synthetic_mel(24, seed=5)
    mel_n = (mel / hparams.sp.max_abs_value).astype(np.float32)
    with torch.no_grad():
        mp = model.pad_tensor(torch.from_numpy(mel_n[None]).transpose(1, 2), pad=model.pad, side="both")
        m_up, a_up = model.upsample(mp.transpose(1, 2))
    m_up, a_up = m_up[0].numpy(), a_up[0].numpy()
    np.savez_compressed(os.path.join(OUT, "cond_raw9.npz"), mel_T=24, mel_seed=5, w_seed=11,
                        aux_frames=a_up[::200].copy(), mels_sub=m_up[::37].copy(),
                        mels_sha=sha(m_up), aux_sha=sha(a_up))
    wav, logits, fed, samples, _ = run_generate(model, fv, mel_n, seed=3, batched=True, target=1000, overlap=200)
    C = 512
    idx = np.rint((samples + 1.0) * (C - 1) / 2.0).astype(np.int16)
    keep = np.r_[0:12, 700:706, 1394:1400]
    np.savez_compressed(os.path.join(OUT, "gen_raw9_batched.npz"), mel_T=24, mel_seed=5, w_seed=11, seed=3,
                        target=1000, overlap=200, bits=9, index=idx[:, :-1], wav=wav,
                        logit_steps=keep, logits=logits[:, keep].copy())
    print("G1", logits.shape, wav.shape)

    # ---- G3: RAW 9-bit unbatched ----------------------------------------------------------------
    mel3 = weights.synthetic_mel(22, seed=6)
    mel3n = (mel3 / 4.0).astype(np.float32)
    wav, logits, fed, samples, _ = run_generate(model, fv, mel3n, seed=4, batched=False, target=1000, overlap=200)
    idx = np.rint((samples + 1.0) * (C - 1) / 2.0).astype(np.int16)
    keep = np.r_[0:8, 2200:2204, 4396:4400]
    np.savez_compressed(os.path.join(OUT, "gen_raw9_unbatched.npz"), mel_T=22, mel_seed=6, w_seed=11, seed=4,
                        bits=9, index=idx[:, :-1], wav=wav, logit_steps=keep, logits=logits[:, keep].copy())
    print("G3", logits.shape, wav.shape)

    # ---- G6: teacher-forced forward() (row a17) -------------------------------------------------
    rng = np.random.default_rng(9)
    xs = rng.uniform(-1, 1, size=(1, 22 * 200)).astype(np.float32)
    with torch.no_grad():
        mp = np.zeros((1, 80, 26), np.float32)
        mp[0, :, 2:24] = mel3n
        tf = model(torch.from_numpy(xs), torch.from_numpy(mp))[0].numpy()
    model.step.data.zero_()
    keep = np.r_[0:6, 1000:1003, 4397:4400]
    np.savez_compressed(os.path.join(OUT, "teacher_forced_raw9.npz"), mel_T=22, mel_seed=6, w_seed=11, x_seed=9,
                        logit_steps=keep, logits=tf[keep].copy(), logits_sha=sha(tf))
    print("G6", tf.shape)

    # ---- G2: MOL batched ------------------------------------------------------------------------
    sdm = weights.make_state_dict(seed=12, bits=9, mode="MOL")
    modelm = build_reference_model(base, hparams, sdm, 9, "MOL")
    wav, logits, fed, samples, mix = run_generate(modelm, fv, mel_n, seed=7, batched=True, target=1000, overlap=200)
    np.savez_compressed(os.path.join(OUT, "gen_mol_batched.npz"), mel_T=24, mel_seed=5, w_seed=12, seed=7,
                        target=1000, overlap=200, samples=samples[:, :-1], mix=mix.astype(np.int8), wav=wav,
                        logits_sub=logits[:, ::8].copy())
    print("G2", logits.shape, wav.shape)

    # ---- G4: fold / unfold index arithmetic (a8, a13) --------------------------------------------
    cases = [(4800, 1000, 200), (4400, 1000, 200), (3800, 1000, 200), (1000, 1000, 200), (5000, 700, 151),
             (9000, 3000, 1500), (2400, 600, 1), (7777, 1234, 321), (160000, 8000, 800), (12345, 50, 7)]
    fold = {}
    for n, (N, tg, ov) in enumerate(cases):
        ramp = torch.arange(N * 2, dtype=torch.float32).reshape(1, N, 2)
        f = model.fold_with_overlap(ramp, tg, ov).numpy()
        y = np.random.default_rng(100 + n).uniform(-1, 1, size=f.shape[:2])
        un = model.xfade_and_unfold(y.copy(), tg, ov)
        fold["case%d" % n] = np.array([N, tg, ov, f.shape[0], f.shape[1]], np.int64)
        fold["fold_first%d" % n] = f[:, 0, 0].astype(np.int64)           # ramp value at each fold start
        fold["fold_last%d" % n] = f[:, -1, 1].astype(np.int64)           # 0 where zero padded
        fold["fold_sha%d" % n] = np.frombuffer(bytes.fromhex(sha(f)), np.uint8)
        fold["unfold_sha%d" % n] = np.frombuffer(bytes.fromhex(sha(un)), np.uint8)
        if N <= 9000:
            fold["unfold%d" % n] = un
    np.savez_compressed(os.path.join(OUT, "fold_unfold.npz"), n_cases=len(cases), **fold)

    # ---- G5: post chain (a14, a15, a16) ----------------------------------------------------------
    rng = np.random.default_rng(21)
    y = rng.uniform(-1, 1, size=6000)
    np.savez_compressed(os.path.join(OUT, "post_chain.npz"), y=y,
                        mu512=ref_audio.decode_mu_law(y, 512, False), mu1024=ref_audio.decode_mu_law(y, 1024, False),
                        deemph=ref_audio.de_emphasis(y),
                        labels=ref_audio.label_2_float(np.arange(512, dtype=np.float64), 9))

    # ---- G7: pruning mask + libwavernn compression (a18, a19) ------------------------------------
    rng = np.random.default_rng(31)
    W = rng.standard_normal((24, 32)).astype(np.float32)
    W3 = rng.standard_normal((36, 16)).astype(np.float32)

    class _L:  # minimal stand-ins so PruneMask.__init__ can run (pruner.py:12-46)
        pass
    lin = torch.nn.Linear(32, 24, bias=True)
    gru = torch.nn.GRU(16, 12, batch_first=True)
    pm_lin = PruneMask(lin, True)
    pm_gru = PruneMask(gru, True)
    m_lin = pm_lin.mask_from_matrix(torch.from_numpy(W), 0.9, 4).numpy()
    m_gru = pm_gru.mask_from_matrix(torch.from_numpy(W3), 0.75, 4).numpy()
    hpz = copy.deepcopy(hparams.wavernn_fatchord)
    wts, idx = ref_convert.compress(W * m_lin, hpz)
    np.savez_compressed(os.path.join(OUT, "prune_compress.npz"), W=W, W3=W3, mask_lin=m_lin, mask_gru=m_gru,
                        comp_w=wts, comp_idx=idx)
    print("done ->", OUT)


if __name__ == "__main__":
    main()
