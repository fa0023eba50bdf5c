"""Mint golden vectors for the runtimeracer topology from the UNMODIFIED reference (build container only).
TEST INFRASTRUCTURE.  `python -m oracle.make_golden_rr` -> tests/golden/rr_{raw9,mol}.npz: generate() traces (logits, fed-back
samples, float64 wav) under the injected Philox noise, and a teacher-forced forward().  Same method as oracle/make_golden.py."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import runtimeracer_oracle as rr, weights  # noqa: E402
from oracle.make_golden import NoiseInjector, OUT  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402


def build(base, hparams, sd, bits, mode):
    hp = copy.deepcopy(hparams.wavernn_runtimeracer)
    hp.bits, hp.mode = bits, mode
    model, _ = base.init_voc_model(base.MODEL_TYPE_RUNTIMERACER, torch.device("cpu"), override_hp_runtimeracer=hp)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return model.eval()


def run_generate(model, mod, mel_norm, seed, batched, target, overlap):
    logits, fed = [], []
    h1 = model.fc5.register_forward_hook(lambda m, i, o: logits.append(o.detach().numpy().copy()))
    h2 = model.I.register_forward_hook(lambda m, i, o: fed.append(i[0][:, 0].detach().numpy().copy()))
    try:
        with NoiseInjector(mod, seed):
            wav = model.generate(torch.from_numpy(mel_norm[None]), batched, target, overlap, True, True, progress_callback=lambda *a: None)
    finally:
        h1.remove()
        h2.remove()
    model.eval()
    logits = np.stack(logits, axis=1)
    fed = np.stack(fed, axis=1)
    samples = np.concatenate([fed[:, 1:], np.zeros((fed.shape[0], 1), np.float32)], axis=1)
    return wav, logits, samples


def main():
    base, _fv, hparams, _ = import_reference()
    from vocoder.models import runtimeracer_version as rv
    T, tg, ov = 24, 1000, 200
    mel = weights.synthetic_mel(T, seed=31) / np.float32(4.0)
    for mode, seed in (("RAW", 21), ("MOL", 22)):
        sd = rr.make_state_dict_rr(seed=seed, bits=9, mode=mode)
        model = build(base, hparams, sd, 9, mode)
        wav, logits, samples = run_generate(model, rv, mel, 5, True, tg, ov)
        # teacher-forced forward() on the (already padded) mel and the first fold's fed-back samples (runtimeracer_version.py:136-196)
        mp = np.pad(mel, ((0, 0), (2, 2)))
        x = np.zeros((1, T * 200), np.float32)
        n = min(samples.shape[1], T * 200)
        x[0, 1:n] = samples[0, :n - 1]
        with torch.no_grad():
            tf = model(torch.from_numpy(x), torch.from_numpy(mp[None])).numpy()
        np.savez_compressed(os.path.join(OUT, "rr_%s.npz" % ("raw9" if mode == "RAW" else "mol")), mel=mel, seed=np.int64(5),
                            wseed=np.int64(seed), target=np.int64(tg), overlap=np.int64(ov), wav=wav,
                            logits=logits[:, :48].astype(np.float32), samples=samples, tf_x=x[:, :120], tf_logits=tf[0, :120].astype(np.float32))
        print(mode, "folds", logits.shape[0], "steps", logits.shape[1], "wav", wav.shape)


if __name__ == "__main__":
    main()
