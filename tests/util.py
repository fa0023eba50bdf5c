import os

import numpy as np

from oracle import weights


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def make_model(seed=11, bits=9, mode="RAW", device=0, prune=None):
    """B200 engine loaded with the deterministic oracle weights."""
    import copy
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_fatchord)
    hp.bits, hp.mode = bits, mode
    sd = weights.make_state_dict(seed=seed, bits=bits, mode=mode)
    if prune:
        sd = weights.prune_state_dict(sd, z=prune)
    model, _ = base.init_voc_model(base.MODEL_TYPE_FATCHORD, device, override_hp_fatchord=hp)
    model.load_state_dict(sd)
    return model, sd


def golden(name):
    here = os.path.dirname(os.path.abspath(__file__))
    return np.load(os.path.join(here, "golden", name))


def norm_mel(T, seed):
    return weights.synthetic_mel(T, seed=seed) / np.float32(4.0)
