#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 WaveRNN vocoder engine (contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W [--workload cfg3] [--precision f32|f16]
    python bench.py --impl reference ...      # CPU arm: torch-CPU port of the reference loop (oracle/torch_port.py)

One "step" = one pass of the hot path (infer_waveform: conditioning -> fold -> sample loop -> xfade ->
mu-law -> de-emphasis) over one batch of synthetic mel input.  Metric = BASELINE.json's: vocoder output
samples/sec (x real-time = / 16000).  N GPUs: every rank vocodes its own utterance(s) (weak scaling, no
collective on the data path; torch.distributed/NCCL is used only for the timing barrier and max).
  value : inputs already resident in HBM, output left in HBM, timed with CUDA events on the engine's stream
  e2e   : the public call rtvc_b200.vocoder.inference.infer_waveform with HOST buffers (pinned), copies timed
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mode, bits, seconds, batched, target, overlap)  -- BASELINE.json configs / SURVEY.md section 8(d)
    "cfg1": ("RAW", 9, 10, True, 8000, 800),      # 19 folds x 9600 steps
    "cfg2": ("RAW", 9, 3, False, 0, 0),           # 1 x 48000 steps (per-step latency)
    # cfg3: the config the target is quoted on ("batched-fold 60 s utterance, 128+ folds, MOL").  The fold plan is the
    # caller's choice (infer_waveform's target / overlap arguments); this one keeps the reference's 10:1 target:overlap ratio
    # (vocoder defaults 8000/800) and fills the loop kernel's 1024 fold slots (2 groups x 4 pipelined sets x 128).
    # cfg3ref is the plan infer_waveform(mel) picks when the caller passes nothing: gen_target=3000 / gen_overlap=1500
    # (config/hparams.py:283-284) -> 213 folds x 6000 steps; cfg3a is SURVEY.md's alternative 137 folds x 8000 steps.
    "cfg3": ("MOL", 9, 60, True, 853, 85),
    "cfg3ref": ("MOL", 9, 60, True, 3000, 1500),
    "cfg3a": ("MOL", 9, 60, True, 6000, 1000),
    "cfg1x60": ("RAW", 9, 60, True, 6000, 1000),  # RAW at the cfg3 shape
    # cfg4: cfg1 with weights pruned to ~90 % zero 1x4 groups (vocoder/pruner.py); CPU comparator = libwavernn port
    "cfg4": ("RAW", 9, 10, True, 8000, 800),
    # cfg5: 256 utterances of 5..20 s (numpy default_rng(2)), all folds pooled, sharded by utterance over the ranks
    "cfg5": ("RAW", 9, 0, True, 6000, 1000),
}
PRUNED = {"cfg4"}
MULTI = {"cfg5"}


def workload_mels(wl, rank, world):
    """Synthetic mels (synthesizer range [-4,4]) of this rank and the global index of its first utterance."""
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import synth as weights
    mode, bits, seconds, batched, target, overlap = WORKLOADS[wl]
    if wl in MULTI:
        lens = np.random.default_rng(2).integers(5, 21, size=256)
        idx = [i for i in range(256) if i % world == rank]
        return [weights.synthetic_mel(80 * int(lens[i]), seed=100 + i) for i in idx], idx
    return [weights.synthetic_mel(80 * seconds, seed=1 + rank)], [rank]


PEAKS_FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def macs_per_row_step(C):
    """Algorithmic MACs of the REFERENCE's step per fold: I + rnn1 + rnn2 + fc1 + fc2 + fc3."""
    return 57344 + 1572864 + 1622016 + 278528 + 278528 + 512 * C


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p, "measured"
    except Exception:
        return PEAKS_FALLBACK, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(wl, sd, mode, mel_norm, batched, target, overlap, seconds):
    """CPU comparator on the host cores, bounded sample.  cfg4: the libwavernn C++ port (block-sparse engine, one
    engine per thread as vocoder/libwavernn/inference.py:43-54); otherwise the torch-CPU port of the reference loop."""
    if wl in PRUNED:
        import tempfile
        from oracle import libwavernn_io
        from oracle.libwavernn_runner import build, time_threads
        build(force=True)                                   # -march=native: rebuild on the box it runs on
        path = os.path.join(tempfile.mkdtemp(), "model.bin")
        libwavernn_io.write_bin(path, sd)
        threads = os.cpu_count()
        n, dt = time_threads(path, mel_norm, threads, 4)                     # calibration
        frames = int(max(4, min(mel_norm.shape[1] - 1, 4 * seconds / max(dt, 1e-3))))
        n, dt = time_threads(path, mel_norm, threads, frames)
        return {"value": n / dt, "unit": "samples/s", "cores": threads, "kind": "port",
                "sample": "libwavernn C++ port (oracle/libwavernn_port.cpp, -O2 -ffast-math -march=native), %d threads x %d frames "
                          "each of the pruned model; steady-state samples/s (fold overlap not counted)" % (threads, frames)}
    from oracle.torch_port import time_generate
    r = time_generate(sd, mode, mel_norm, batched, target, overlap, threads=os.cpu_count(), time_budget_s=seconds, seed=1)
    return {"value": r["out_samples"] / r["est_total_seconds"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
            "sample": "%d of %d loop steps x %d folds (torch CPU port of the reference loop), conditioning in full; "
                      "extrapolated linearly" % (r["steps_done"], r["steps_total"], r["folds"])}


def reference_arm(args, wl):
    """CPU arm (rank 0 only): the reference's own CPU implementation of the path on the box's host cores -- the
    torch-CPU port of its loop (or, for the pruned workload, the libwavernn C++ port) -- on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import synth as weights
    mode, bits, seconds, batched, target, overlap = WORKLOADS[wl]
    sd = weights.make_state_dict(seed=0, bits=bits, mode=mode)
    if wl in PRUNED:
        sd = weights.prune_state_dict(sd, z=0.9)
    mel = workload_mels(wl, 0, 1)[0][0] / np.float32(4.0)
    budget = float(args.cpu_seconds) / max(1, args.steps + args.warmup)
    vals = [cpu_baseline(wl, sd, mode, mel, batched, target, overlap, budget) for _ in range(args.warmup + args.steps)][args.warmup:]
    value = float(np.mean([v["value"] for v in vals]))
    out_samples = (mel.shape[1] - 1) * 200
    line = {
        "impl": "reference", "metric": "vocoder_output_samples_per_sec", "value": value, "unit": "samples/s",
        "x_realtime": value / 16000.0, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": out_samples / value * 1e3, "higher_is_better": True,
        "scaling": "strong" if wl in MULTI else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl),
        "cpu_baseline": dict(vals[-1], value=value),
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def workload_config(wl):
    mode, bits, seconds, batched, target, overlap = WORKLOADS[wl]
    return {"workload": "%s: WaveRNN fatchord %s%s, %s synthetic 80-mel @16 kHz, %s" % (
        wl, mode, (" %d-bit" % bits) if mode == "RAW" else "",
        "256 utterances of 5-20 s" if wl in MULTI else "%d s" % seconds,
        ("batched target=%d overlap=%d" % (target, overlap)) if batched else "unbatched (single fold)"),
        "weights": "random-init rnn_dims=512 fc_dims=512 hop=200", "cache": "L2 flushed (256 MiB write) between timed steps"}


_REAL_STDOUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout: native libraries (e.g. NCCL's version banner) write to fd 1 too, so the
    run happens with fd 1 pointing at stderr and the line goes to the saved descriptor at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["f32", "f16", "sparse"],
                    help="f16: tensor-core loop (fp16 operands, fp32 accumulate/state); f32: parity-mode loop")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU time budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _guard_stdout()
    wl = args.workload
    if args.precision is None:      # defaults = what the facade's PREC_AUTO picks (fatchord_version.resolve_precision): pruned model ->
        # block-sparse cluster loop; fewer than 24 folds in the call (cfg1: 19, cfg2: 1) -> fp32 loop; else the tensor-core loop
        args.precision = "sparse" if wl in PRUNED else ("f32" if (not WORKLOADS[wl][3] or wl == "cfg1") else "f16")
    if args.impl == "reference":
        return reference_arm(args, wl)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as entry
    entry.build()
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    from rtvc_b200.config import hparams
    from rtvc_b200.vocoder import inference
    from rtvc_b200 import synth as weights      # deterministic synthetic weights / mels (input generation only)

    mode, bits, seconds, batched, target, overlap = WORKLOADS[wl]
    hp = copy.deepcopy(hparams.wavernn_fatchord)
    hp.bits, hp.mode = bits, mode
    hparams.wavernn_fatchord.bits, hparams.wavernn_fatchord.mode = bits, mode   # infer_waveform reads the globals
    sd = weights.make_state_dict(seed=0, bits=bits, mode=mode)
    if wl in PRUNED:
        sd = weights.prune_state_dict(sd, z=0.9)
    model = inference.load_state(sd, devices=[local_rank], override_hp_fatchord=hp)
    model.precision = {"f32": _native.PREC_F32, "f16": _native.PREC_F16, "sparse": _native.PREC_SPARSE_F32}[args.precision]
    mels_raw, utt_idx = workload_mels(wl, rank, world)                      # synthesizer range [-4, 4]
    mels_host = [torch.from_numpy(m).pin_memory() for m in mels_raw]
    mels_dev = [(m / 4.0).cuda() for m in mels_host]
    out_samples = sum((m.shape[1] - 1) * 200 for m in mels_raw)           # of this rank
    wav_dev = torch.empty(out_samples, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    C = model.n_classes
    import ctypes as Ct

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.fill_(1)
        torch.cuda.synchronize()
        rq, arrs, wav, offsets, keep = model._request([np.zeros(m.shape, np.float32) for m in mels_raw], batched, target,
                                                      overlap, hp.mu_law, True, None, want_wav=False, utt_index0=utt_idx[0])
        ptr = (Ct.c_void_p * len(mels_dev))(*[m.data_ptr() for m in mels_dev])
        rq.mels = Ct.cast(ptr, Ct.POINTER(Ct.c_void_p))
        rq.mels_on_device = 1
        rq.wav = wav_dev.data_ptr()
        rq.wav_capacity = out_samples
        rq.wav_on_device = 1
        model._run(rq)
        t = model.last_timings
        return t["ms_h2d"] + t["ms_cond"] + t["ms_loop"] + t["ms_post"] + t["ms_d2h"], dict(t)

    def step_e2e():
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if len(mels_host) == 1:
            wavs = [inference.infer_waveform(mels_host[0].numpy(), normalize=True, batched=batched, target=target, overlap=overlap)]
        else:
            wavs = inference.infer_waveforms([m.numpy() for m in mels_host], normalize=True, batched=batched, target=target,
                                             overlap=overlap, utt_index0=utt_idx[0])
        dt = time.perf_counter() - t0
        assert sum(w.shape[0] for w in wavs) == out_samples
        return dt * 1e3

    sampler = ClockSampler(local_rank)
    # ---- value: HBM-resident ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler.start()
    launches0 = model.launch_count
    per_step, loop_ms, last_t = [], [], None
    for _ in range(args.steps):
        ms, last_t = step_resident()
        per_step.append(ms)
        loop_ms.append(last_t["ms_loop"])
    barrier()
    launches = model.launch_count - launches0
    total_ms = float(sum(per_step))
    # ---- e2e: public API, host buffers ----------------------------------------------------------------------
    step_e2e()
    barrier()
    e2e_ms = [step_e2e() for _ in range(args.steps)]
    barrier()
    clocks = sampler.finish()
    e2e_total = float(sum(e2e_ms))
    all_samples = out_samples
    if world > 1:
        t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0]), float(t[1])
        n = torch.tensor([out_samples], dtype=torch.float64, device="cuda")
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        all_samples = int(n[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = all_samples * args.steps / (total_ms / 1e3)
    e2e_value = all_samples * args.steps / (e2e_total / 1e3)
    pk, pk_kind = peaks()
    F, S = last_t["n_folds"], last_t["n_steps"]
    flops = 2.0 * macs_per_row_step(C) * F * S
    loop_s = float(np.mean(loop_ms)) / 1e3
    achieved = flops / loop_s / 1e12
    peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get("%s:%s" % (wl, args.precision))
    except Exception:
        pass
    floor = model.barrier_floor(20000)
    floor["cluster_us"] = model.cluster_floor(16, 20000)
    n_exch = 5 if (mode == "MOL" or args.precision == "sparse") else 6       # h1, h2, f1, f2, (logits,) x per step
    floor_us = {"f32": floor["ll_us"], "f16": floor["counter_us"], "sparse": floor["cluster_us"]}[args.precision]
    line = {
        "metric": "vocoder_output_samples_per_sec", "value": value, "unit": "samples/s", "x_realtime": value / 16000.0,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if wl in MULTI else "weak", "vs_baseline": None,
        "dtype": "f16" if args.precision == "f16" else "f32", "data": "synthetic",
        "precision_note": ("fp32 weights/FMA/state (parity mode)" if args.precision == "f32" else
                           "fp32 block-sparse (1x4 groups) cluster-local loop" if args.precision == "sparse" else
                           "fp16 weights+activations on tcgen05, fp32 accumulate, fp32 recurrent state and conditioning; "
                           "teacher-forced logits 4.2e-4 rel, 100% identical draws on the golden run (tests/test_gpu_tc.py)"),
        "config": dict(workload_config(wl), folds=F, loop_steps=S, per_gpu=("256 utterances sharded by utterance" if wl in MULTI else "one utterance per GPU, independent"),
                       pruned=("~90% zero 1x4 groups (pruner.py rule), run through the dense kernels" if wl in PRUNED else None)),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "x_realtime": e2e_value / 16000.0,
                "h2d_bytes_per_step": int(sum(m.nbytes for m in mels_raw)), "d2h_bytes_per_step": int(out_samples * 8),
                "ms_per_step": e2e_total / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("bytes"), "traffic_source": (traffic or {}).get("source"), "kernel": {"f32": "wrnn_loop_f32_kernel", "f16": "wrnn_loop_tc_kernel", "sparse": "wrnn_loop_sparse_kernel"}[args.precision],
                     "peak_source": pk_kind + " bf16_tflops_sustained", "kernel_ms": loop_s * 1e3,
                     "algorithmic_flops_per_launch": flops},
        "loop": {"us_per_step": loop_s * 1e6 / (S * max(1, last_t["n_launches"])), "fold_sets_per_group": (max(1, min(4, -(-F // 256))) if args.precision == "f16" else None),
                 "exchanges_per_step": n_exch, "exchange_floor_us": floor["ll_us"],
                 "counter_barrier_floor_us": floor["counter_us"], "cluster16_exchange_floor_us": floor["cluster_us"],
                 "step_over_floor": (loop_s * 1e6 / (S * max(1, last_t["n_launches"]))) / (n_exch * floor_us),
                 "floor_used": {"f32": "flag-in-data exchange through L2", "f16": "fence+atomic counter barrier through L2",
                                "sparse": "DSMEM stores + cluster barrier (16 CTAs)"}[args.precision]},
        "phases_ms": {k: last_t[k] for k in ("ms_h2d", "ms_cond", "ms_loop", "ms_post", "ms_d2h")},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl, sd, mode, mels_raw[0] / np.float32(4.0), batched, target, overlap, args.cpu_seconds)
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
