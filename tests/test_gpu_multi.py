"""Multi-GPU fold sharding inside one process (skipped with < 2 GPUs): identical waveform to one GPU."""
import numpy as np
import pytest

from tests.util import norm_mel

pytestmark = pytest.mark.gpu


def test_two_engines_match_one(monkeypatch):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import copy
    import rtvc_b200  # noqa: F401
    from rtvc_b200.config import hparams
    from rtvc_b200.vocoder import inference
    from oracle import weights
    hp = copy.deepcopy(hparams.wavernn_fatchord)
    hp.bits = 9
    hparams.wavernn_fatchord.bits = 9
    sd = weights.make_state_dict(seed=11, bits=9, mode="RAW")
    mel = norm_mel(60, 2) * 4.0
    inference.load_state(sd, devices=[0], override_hp_fatchord=hp)
    inference.set_seed(5)
    one = inference.infer_waveform(mel, target=800, overlap=100)
    inference.load_state(sd, devices=[0, 1], override_hp_fatchord=hp)
    inference.set_seed(5)
    two = inference.infer_waveform(mel, target=800, overlap=100)       # 13 folds: the facade keeps one GPU (SHARD_MIN_FOLDS)
    assert np.array_equal(one, two)
    monkeypatch.setattr(inference, "SHARD_MIN_FOLDS", 0)               # force the fold-range split over both engines
    inference.set_seed(5)
    split = inference.infer_waveform(mel, target=800, overlap=100)
    assert np.array_equal(one, split)
    inference.set_seed(5)
    many = inference.infer_waveforms([mel, mel[:, :40], mel[:, :33]], target=800, overlap=100)
    assert [len(w) for w in many] == [59 * 200, 39 * 200, 32 * 200]
    inference.unload()
