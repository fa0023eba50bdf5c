"""The geneing topology on the GPU (csrc/loop_gn.cu + engine_gn.inc) against vectors minted from
the UNMODIFIED reference (tests/golden/gn_bits9.npz, oracle/make_golden_gn.py) and against the oracle
(oracle/geneing_oracle.py).  Gates: teacher-forced logits within 1e-4 relative (fp32 loop), >= 99.9 % identical draws, float64
wav within 1e-3."""
import copy
import os

import numpy as np
import pytest

from oracle import geneing_oracle as gn
from tests.util import norm_mel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def make_gn(seed, bits=9, device=0):
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_geneing)
    hp.bits, hp.mode = bits, "BITS"
    sd = gn.make_state_dict_gn(seed=seed, bits=bits)
    model, _ = base.init_voc_model(base.MODEL_TYPE_GENEING, device, override_hp_geneing=hp)
    model.load_state_dict(sd)
    assert base.get_model_type(model) == base.MODEL_TYPE_GENEING and model.mode == "BITS" and model.aux_dims == 32
    return model, sd


def test_geneing_gpu_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "gn_bits9.npz"))
    model, sd = make_gn(int(g["wseed"]), 9)
    tg, ov = int(g["target"]), int(g["overlap"])
    S = tg + 2 * ov
    ref = g["samples"]
    forced = np.zeros((ref.shape[0], S), np.float32)
    forced[:, :ref.shape[1]] = ref
    n = g["logits"].shape[1]
    o = model.generate_debug(g["mel"], True, tg, ov, forced=forced, want_logits=True, seed=int(g["seed"]), max_steps=n)
    assert dict(model.last_timings)["loop_kernel"] == "wrnn_loop_gn_kernel"
    err = _rel(o["logits"], g["logits"])
    f = model.generate_debug(g["mel"], True, tg, ov, seed=int(g["seed"]))
    k = ref.shape[1] - 1
    agree = float((f["samples"][:, :k] == ref[:, :k]).mean())
    wav = model.generate(g["mel"][None], True, tg, ov, False, True, seed=int(g["seed"]))      # mu_law False: config/hparams.py:292
    werr = float(np.abs(wav - g["wav"]).max())
    print("geneing BITS vs reference golden: logits rel err %.3e, draw agreement %.5f, wav max err %.3e" % (err, agree, werr))
    assert err < 1e-4, err
    assert agree >= 0.999, agree
    assert wav.shape == g["wav"].shape and wav.dtype == np.float64
    assert werr < 1e-3, werr


def test_geneing_gpu_waves_and_unbatched_vs_oracle():
    """More folds than one launch holds (124 folds = two waves of <= 96) and the unbatched path: first steps against the oracle,
    teacher-forced on the kernel's own samples."""
    model, sd = make_gn(7, 9)
    for batched, T, tg, ov, steps in [(True, 496, 700, 100, 24), (False, 12, 0, 0, 64)]:
        mel = norm_mel(T, 4)
        o = model.generate_debug(mel, batched, tg, ov, want_logits=True, seed=7, max_steps=steps)
        F = o["samples"].shape[0]
        S = tg + 2 * ov if batched else T * 200
        forced = np.zeros((F, S), np.float32)
        forced[:, :steps] = o["samples"]
        t = gn.generate_gn(mel, sd, 7, bits=9, batched=batched, target=tg, overlap=ov, forced=forced, max_steps=steps)
        err = _rel(o["logits"], t["logits"])
        agree = float((o["samples"] == t["samples"]).mean())
        print("geneing %s, %d folds x %d steps vs oracle: logits rel err %.3e, draw agreement %.5f" % ("batched" if batched else "unbatched", F, steps, err, agree))
        assert (F > 96) == batched
        assert err < 1e-4 and agree >= 0.999, (err, agree)


def test_geneing_rejects_beta_mode():
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_geneing)
    hp.mode = "RAW"
    with pytest.raises(NotImplementedError):
        base.init_voc_model(base.MODEL_TYPE_GENEING, 0, override_hp_geneing=hp)


def test_geneing_mol_vs_oracle():
    """Mode 'MOL' of this topology (geneing_version.py:217-223: the fatchord mixture rule on fc3's 30 outputs): first 48 steps against
    the oracle, teacher-forced on the kernel's own samples (no reference-minted vector exists for this mode: the golden is 'BITS')."""
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder.models import base
    from rtvc_b200.config import hparams
    hp = copy.deepcopy(hparams.wavernn_geneing)
    hp.mode = "MOL"
    sd = gn.make_state_dict_gn(seed=9, bits=9, mode="MOL")
    model, _ = base.init_voc_model(base.MODEL_TYPE_GENEING, 0, override_hp_geneing=hp)
    model.load_state_dict(sd)
    assert model.n_classes == 30
    mel = norm_mel(40, 6)
    steps, tg, ov = 48, 600, 100
    o = model.generate_debug(mel, True, tg, ov, want_logits=True, seed=11, max_steps=steps)
    F = o["samples"].shape[0]
    forced = np.zeros((F, tg + 2 * ov), np.float32)
    forced[:, :steps] = o["samples"]
    t = gn.generate_gn(mel, sd, 11, bits=9, batched=True, target=tg, overlap=ov, forced=forced, max_steps=steps, mode="MOL")
    err = _rel(o["logits"], t["logits"])
    agree = float((np.abs(o["samples"] - t["samples"]) < 1e-4).mean())
    print("geneing MOL, %d folds x %d steps vs oracle: logits rel err %.3e, draw agreement %.5f" % (F, steps, err, agree))
    assert err < 1e-4 and agree >= 0.999, (err, agree)
