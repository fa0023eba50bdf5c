"""Times the UNMODIFIED reference (vocoder.models.base.init_voc_model(...).generate(...), fatchord_version.py:155-259) on
the host cores, from the copy under baseline/_ref (oracle/install_ref.py).  BENCH INFRASTRUCTURE ONLY (bench.py --impl
reference and its cpu_baseline leg).

The reference's generate() has no step limit; a bounded sample is taken through its own progress_callback contract
(:234-236, called every 100 steps): the callback stamps the clock and, once the time budget is spent, raises to leave the
loop.  Reported: steady-state seconds per step (between the first and the last callback), the measured conditioning +
fold time before the first step, and the whole-call estimate  t_before + per_step * S  (the numpy post chain -- ~30 ms for
60 s, BASELINE.md section 2 -- is not added).  CUDA must be hidden BEFORE torch is imported (the reference moves its
tensors to the GPU whenever one is visible, fatchord_version.py:166-168)."""
import copy
import os
import time

import numpy as np


class _Stop(Exception):
    pass


def available():
    from . import ref_import
    return ref_import.find_reference_root() is not None


def time_reference(sd, mode, bits, mel_norm, batched, target, overlap, threads=None, budget_s=10.0, model_type="fatchord-wavernn"):
    """sd: numpy state dict (oracle/weights.py); mel_norm: (80, T) float32 already divided by max_abs_value."""
    assert os.environ.get("CUDA_VISIBLE_DEVICES", None) == "", "hide the GPUs before importing torch"
    import torch
    from . import ref_import
    base, fv, hpm, _ = ref_import.import_reference()
    torch.set_num_threads(int(threads or os.cpu_count()))
    hp = copy.deepcopy({"fatchord-wavernn": hpm.wavernn_fatchord, "runtimeracer-wavernn": hpm.wavernn_runtimeracer,
                        "geneing-wavernn": hpm.wavernn_geneing}[model_type])
    hp.bits, hp.mode = bits, ("BITS" if model_type == "geneing-wavernn" else mode)
    model, _ = base.init_voc_model(model_type, torch.device("cpu"), override_hp_fatchord=hp, override_hp_runtimeracer=hp, override_hp_geneing=hp)
    state = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
    model.load_state_dict(state, strict=False)
    model.eval()
    stamps = []

    def cb(i, seq_len, b_size, gen_rate):
        now = time.perf_counter()
        stamps.append((i, now, seq_len, b_size))
        if len(stamps) >= 2 and now - stamps[0][1] > budget_s:
            raise _Stop()

    mel_t = torch.from_numpy(np.ascontiguousarray(mel_norm))[None]
    t_start = time.perf_counter()
    finished = False
    try:
        model.generate(mel_t, batched, target, overlap, hp.mu_law, hpm.sp.preemphasize, cb)
        finished = True
    except _Stop:
        pass
    t_end = time.perf_counter()
    out_samples = (mel_norm.shape[1] - 1) * hpm.sp.hop_size
    i0, t0, S, B = stamps[0]
    i1, t1 = stamps[-1][0], stamps[-1][1]
    if finished:
        total = t_end - t_start
        per_step = (t1 - t0) / max(1, i1 - i0)
    else:
        per_step = (t1 - t0) / max(1, i1 - i0)
        total = (t0 - t_start) + per_step * S
    return dict(value=out_samples / total, seconds_total=total, seconds_before_loop=t0 - t_start, us_per_step=per_step * 1e6,
                steps_measured=int(i1 - i0), steps_total=int(S), folds=int(B), threads=int(torch.get_num_threads()),
                extrapolated=not finished, measured_fraction=float(min(1.0, (i1 - i0) / max(1, S))))
