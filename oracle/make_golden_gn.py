"""Mint golden vectors for the geneing topology (mode 'BITS') from the UNMODIFIED reference (build container only).
TEST INFRASTRUCTURE.  `python -m oracle.make_golden_gn` -> tests/golden/gn_bits9.npz.  Same method as oracle/make_golden.py."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import geneing_oracle as gn  # noqa: E402
from oracle.make_golden import NoiseInjector, OUT  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from oracle.weights import synthetic_mel  # noqa: E402


def main():
    base, _fv, hparams, _ = import_reference()
    from vocoder.models import geneing_version as gv
    hp = copy.deepcopy(hparams.wavernn_geneing)
    hp.bits, hp.mode = 9, "BITS"
    sd = gn.make_state_dict_gn(seed=41, bits=9)
    model, _ = base.init_voc_model(base.MODEL_TYPE_GENEING, torch.device("cpu"), override_hp_geneing=hp)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model.eval()
    T, tg, ov = 24, 1000, 200
    mel = synthetic_mel(T, seed=33) / np.float32(4.0)
    logits, fed = [], []
    h1 = model.fc3.register_forward_hook(lambda m, i, o: logits.append(o.detach().numpy().copy()))
    h2 = model.I.register_forward_hook(lambda m, i, o: fed.append(i[0][:, 0].detach().numpy().copy()))
    with NoiseInjector(gv, 5):
        wav = model.generate(torch.from_numpy(mel[None]), True, tg, ov, hp.mu_law, True, progress_callback=lambda *a: None)
    h1.remove(); h2.remove()
    logits = np.stack(logits, axis=1)
    fed = np.stack(fed, axis=1)
    samples = np.concatenate([fed[:, 1:], np.zeros((fed.shape[0], 1), np.float32)], axis=1)
    model.eval()                                   # generate() leaves train mode on (Q1)
    with torch.no_grad():
        m_up, a_up = model.upsample(torch.from_numpy(np.pad(mel, ((0, 0), (2, 2)))[None]))
    np.savez_compressed(os.path.join(OUT, "gn_bits9.npz"), mel=mel, seed=np.int64(5), wseed=np.int64(41), target=np.int64(tg),
                        overlap=np.int64(ov), wav=wav, logits=logits[:, :48].astype(np.float32), samples=samples,
                        up_mels=m_up[0, ::37].numpy(), up_aux=a_up[0, ::37].numpy())
    print("folds", logits.shape[0], "steps", logits.shape[1], "wav", wav.shape)


if __name__ == "__main__":
    main()
