#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 WaveRNN vocoder engine (contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W [--workload cfg3ref] [--precision f32|f16]
    python bench.py --impl reference ...      # CPU arm: the UNMODIFIED reference generate() from baseline/_ref on the host cores

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
    # the other two topologies of the reference at the cfg1 shape; fp32 loops
    "rr1": ("RAW", 9, 10, True, 8000, 800),       # runtimeracer-wavernn: 4 x GRU-256 + 5 FC (wrnn_loop_rr_kernel)
    "gn1": ("RAW", 9, 10, True, 8000, 800),       # geneing-wavernn, mode BITS: GRU-256 + 2 FC (wrnn_loop_gn_kernel)
}
PRUNED = {"cfg4"}
MULTI = {"cfg5"}
TOPO = {"rr1": "runtimeracer-wavernn", "gn1": "geneing-wavernn"}       # default: fatchord-wavernn


def topo_state_dict(weights, wl, bits, mode):
    """Synthetic weights of the workload's topology (rtvc_b200/synth.py)."""
    t = TOPO.get(wl)
    if t == "runtimeracer-wavernn":
        return weights.make_state_dict_rr(seed=0, bits=bits, mode=mode)
    if t == "geneing-wavernn":
        return weights.make_state_dict_gn(seed=0, bits=bits)
    return weights.make_state_dict(seed=0, bits=bits, mode=mode)


def topo_macs(wl, C):
    """Algorithmic MACs of the REFERENCE's step per fold for the workload's topology."""
    t = TOPO.get(wl)
    if t == "runtimeracer-wavernn":      # runtimeracer_version.py:119-131: I, rnn1..4 (rnn3 takes 288), fc1 / fc3 (288 -> 256), fc2 / fc4, fc5
        return 112 * 256 + 3 * 256 * (256 * 3 + 288 + 4 * 256) + 2 * 288 * 256 + 2 * 256 * 256 + 256 * C
    if t == "geneing-wavernn":           # geneing_version.py:107-113: I, rnn1, fc1 (288 -> 128), fc3
        return 112 * 256 + 3 * 256 * 2 * 256 + 288 * 128 + 128 * C
    return macs_per_row_step(C)


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


def libwavernn_port_baseline(sd, mel_norm, seconds, threads=None):
    """The libwavernn C++ restatement (oracle/libwavernn_port.cpp, the reference's flags, one engine per thread as
    vocoder/libwavernn/inference.py:43-54,93-115): steady-state samples/s on the host cores."""
    import tempfile
    from oracle import libwavernn_io
    from oracle.libwavernn_runner import build, time_threads
    build(force=True)                                   # -march=native: rebuild on the box it runs on
    path = os.path.join(tempfile.mkdtemp(), "model.bin")
    libwavernn_io.write_bin(path, sd)
    threads = threads or os.cpu_count()
    n, dt = time_threads(path, mel_norm, threads, 4)                     # calibration
    frames = int(max(4, min(mel_norm.shape[1] - 1, 4 * seconds / max(dt, 1e-3))))
    n, dt = time_threads(path, mel_norm, threads, frames)
    return {"value": n / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": "libwavernn C++ port (oracle/libwavernn_port.cpp, -O2 -ffast-math -march=native), %d threads x %d frames "
                      "each; steady-state samples/s (fold overlap not counted)" % (threads, frames)}


def cpu_baseline(wl, sd, mode, bits, mel_norm, batched, target, overlap, seconds, extras=True):
    """CPU comparator on the host cores, bounded sample (CUDA must be hidden in this process: the reference moves to the GPU
    whenever it sees one).  The UNMODIFIED reference's generate() from baseline/_ref (kind "reference"; oracle/ref_bench.py)
    when that copy travelled, else the torch port of its loop (kind "port").  extras: the same at one thread, and the
    libwavernn C++ port (the reference's other CPU engine); for the pruned workload the libwavernn port is the headline."""
    from oracle import ref_bench
    out = None
    if ref_bench.available():
        topo = TOPO.get(wl, "fatchord-wavernn")
        r = ref_bench.time_reference(sd, mode, bits, mel_norm, batched, target, overlap, threads=os.cpu_count(), budget_s=seconds, model_type=topo)
        out = {"value": r["value"], "unit": "samples/s", "cores": r["threads"], "kind": "reference",
               "extrapolated": r["extrapolated"], "measured_fraction": r["measured_fraction"], "us_per_step": r["us_per_step"],
               "sample": "UNMODIFIED reference base.init_voc_model(...).generate() from baseline/_ref on %d torch threads: conditioning + fold "
                         "timed in full (%.2f s), %d of %d loop steps x %d folds timed through its own progress_callback, the rest "
                         "extrapolated at the measured %.0f us per step" % (r["threads"], r["seconds_before_loop"], r["steps_measured"],
                                                                           r["steps_total"], r["folds"], r["us_per_step"])}
        if extras:
            r1 = ref_bench.time_reference(sd, mode, bits, mel_norm, batched, target, overlap, threads=1, budget_s=max(2.0, seconds / 4), model_type=topo)
            out["one_thread"] = {"value": r1["value"], "us_per_step": r1["us_per_step"], "cores": 1, "measured_fraction": r1["measured_fraction"]}
    else:
        from oracle.torch_port import time_generate
        r = time_generate(sd, mode, mel_norm, batched, target, overlap, threads=os.cpu_count(), time_budget_s=seconds, seed=1)
        out = {"value": r["out_samples"] / r["est_total_seconds"], "unit": "samples/s", "cores": r["threads"], "kind": "port",
               "extrapolated": True, "measured_fraction": r["steps_done"] / max(1, r["steps_total"]),
               "sample": "%d of %d loop steps x %d folds (torch CPU port of the reference loop: baseline/_ref is absent), conditioning in "
                         "full; extrapolated linearly" % (r["steps_done"], r["steps_total"], r["folds"])}
    if (extras and wl not in TOPO) or wl in PRUNED:       # (the libwavernn port restates the fatchord engine only)
        try:
            lw = libwavernn_port_baseline(sd, mel_norm, max(2.0, seconds / 3))
            if wl in PRUNED:
                out = dict(lw, pytorch_reference=out)
            else:
                out["libwavernn_port"] = lw
        except Exception as e:          # (no g++ on the box: keep the line)
            out["libwavernn_port"] = {"unavailable": str(e)[:200]}
    return out


def reference_arm(args, wl):
    """CPU arm (rank 0 only): the reference's own CPU implementation of the path on the box's host cores -- its unmodified
    generate() from baseline/_ref -- on a bounded sample per step (see cpu_baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import synth as weights
    mode, bits, seconds, batched, target, overlap = WORKLOADS[wl]
    sd = topo_state_dict(weights, wl, bits, mode)
    if wl in PRUNED:
        sd = weights.prune_state_dict(sd, z=0.9)
    mel = workload_mels(wl, 0, 1)[0][0] / np.float32(4.0)
    budget = float(args.cpu_seconds) / max(1, args.steps + args.warmup)
    vals = [cpu_baseline(wl, sd, mode, bits, mel, batched, target, overlap, budget, extras=args.cpu_extras and i == args.warmup + args.steps - 1)
            for i in range(args.warmup + args.steps)][args.warmup:]
    value = float(np.mean([v["value"] for v in vals]))
    out_samples = (mel.shape[1] - 1) * 200
    line = {
        "impl": "reference", "metric": "vocoder_output_samples_per_sec", "value": value, "unit": "samples/s",
        "x_realtime": value / 16000.0, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": out_samples / value * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl),
        "cpu_baseline": dict(vals[-1], value=value),
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def workload_config(wl):
    mode, bits, seconds, batched, target, overlap = WORKLOADS[wl]
    topo = TOPO.get(wl, "fatchord-wavernn").split("-")[0]
    dims = {"fatchord": "rnn_dims=512 fc_dims=512", "runtimeracer": "rnn_dims=256 fc_dims=256 (4 GRU, 5 FC)", "geneing": "rnn_dims=256 fc_dims=128 (1 GRU, 2 FC)"}[topo]
    return {"workload": "%s: WaveRNN %s %s%s, %s synthetic 80-mel @16 kHz, %s" % (
        wl, topo, ("BITS" if topo == "geneing" else mode), (" %d-bit" % bits) if mode == "RAW" else "",
        "256 utterances of 5-20 s" if wl in MULTI else "%d s" % seconds,
        ("batched target=%d overlap=%d" % (target, overlap)) if batched else "unbatched (single fold)"),
        "weights": "random-init %s hop=200" % dims, "cache": "L2 flushed (256 MiB write) between timed steps"}


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
    ap.add_argument("--workload", default="cfg3ref", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["f32", "f16", "sparse"],
                    help="f16: tensor-core loops (fp16 operands, fp32 accumulate/state); f32: parity-mode loop")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU time budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra fold plans, the in-run parity check and the sharded runs")
    ap.add_argument("--cpu-extras", type=int, default=1, help="reference arm: also time one thread and the libwavernn port")
    args = ap.parse_args()
    wl = args.workload
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""        # before torch is imported: the reference uses a GPU whenever it sees one
        _guard_stdout()
        return reference_arm(args, wl)
    _guard_stdout()
    if args.precision is None:      # defaults = what the facade's PREC_AUTO picks (fatchord_version.resolve_precision): pruned model ->
        # block-sparse cluster loop; fewer than 8 folds in the call (cfg2: 1) and the two small topologies -> fp32 loops; else the tensor-core loops
        args.precision = "sparse" if wl in PRUNED else ("f32" if (not WORKLOADS[wl][3] or wl in TOPO) else "f16")

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
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None     # host-side rendezvous that leaves the GPUs idle

    import __graft_entry__ as entry
    entry.build()
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    from rtvc_b200.config import hparams
    from rtvc_b200.vocoder import inference
    from rtvc_b200 import synth as weights      # deterministic synthetic weights / mels (input generation only)
    import ctypes as Ct
    PREC = {"f32": _native.PREC_F32, "f16": _native.PREC_F16, "sparse": _native.PREC_SPARSE_F32}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def setup(wl_, precision):
        mode, bits, seconds, batched, target, overlap = WORKLOADS[wl_]
        topo = TOPO.get(wl_, "fatchord-wavernn")
        ghp = {"fatchord-wavernn": hparams.wavernn_fatchord, "runtimeracer-wavernn": hparams.wavernn_runtimeracer, "geneing-wavernn": hparams.wavernn_geneing}[topo]
        hmode = "BITS" if topo == "geneing-wavernn" else mode
        hp = copy.deepcopy(ghp)
        hp.bits, hp.mode = bits, hmode
        ghp.bits, ghp.mode = bits, hmode                                            # infer_waveform reads the globals
        sd = topo_state_dict(weights, wl_, bits, mode)
        if wl_ in PRUNED:
            sd = weights.prune_state_dict(sd, z=0.9)
        model = inference.load_state(sd, topo, devices=[local_rank], override_hp_fatchord=hp, override_hp_runtimeracer=hp, override_hp_geneing=hp)
        model.precision = PREC[precision]
        mels_raw, utt_idx = workload_mels(wl_, rank, world)                      # synthesizer range [-4, 4]
        mels_host = [torch.from_numpy(m).pin_memory() for m in mels_raw]
        mels_dev = [(m / 4.0).cuda() for m in mels_host]
        out_samples = sum((m.shape[1] - 1) * 200 for m in mels_raw)           # of this rank
        wav_dev = torch.empty(out_samples, dtype=torch.float64, device="cuda")
        return dict(wl=wl_, model=model, sd=sd, hp=hp, mels_raw=mels_raw, utt_idx=utt_idx, mels_host=mels_host, mels_dev=mels_dev,
                    out_samples=out_samples, wav_dev=wav_dev, plan=(mode, bits, seconds, batched, target, overlap))

    def step_resident(c):
        mode, bits, seconds, batched, target, overlap = c["plan"]
        model = c["model"]
        flush.fill_(1)
        torch.cuda.synchronize()
        rq, arrs, wav, offsets, keep = model._request([np.zeros(m.shape, np.float32) for m in c["mels_raw"]], batched, target,
                                                      overlap, c["hp"].mu_law, True, None, want_wav=False, utt_index0=c["utt_idx"][0])
        ptr = (Ct.c_void_p * len(c["mels_dev"]))(*[m.data_ptr() for m in c["mels_dev"]])
        rq.mels = Ct.cast(ptr, Ct.POINTER(Ct.c_void_p))
        rq.mels_on_device = 1
        rq.wav = c["wav_dev"].data_ptr()
        rq.wav_capacity = c["out_samples"]
        rq.wav_on_device = 1
        model._run(rq)
        t = model.last_timings
        return t["ms_h2d"] + t["ms_cond"] + t["ms_loop"] + t["ms_post"] + t["ms_d2h"], dict(t)

    def step_e2e(c):
        mode, bits, seconds, batched, target, overlap = c["plan"]
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if len(c["mels_host"]) == 1:
            wavs = [inference.infer_waveform(c["mels_host"][0].numpy(), normalize=True, batched=batched, target=target, overlap=overlap)]
        else:
            wavs = inference.infer_waveforms([m.numpy() for m in c["mels_host"]], normalize=True, batched=batched, target=target,
                                             overlap=overlap, utt_index0=c["utt_idx"][0])
        dt = time.perf_counter() - t0
        assert sum(w.shape[0] for w in wavs) == c["out_samples"]
        return dt * 1e3

    def measure(c, warmup, steps, e2e=True):
        """value leg (HBM-resident) and e2e leg (public API, pinned host buffers) of one workload; max over ranks, summed samples."""
        for _ in range(warmup):
            step_resident(c)
        barrier()
        launches0 = c["model"].launch_count
        per_step, loop_ms, last_t = [], [], None
        for _ in range(steps):
            ms, last_t = step_resident(c)
            per_step.append(ms)
            loop_ms.append(last_t["ms_loop"])
        barrier()
        launches = c["model"].launch_count - launches0
        total_ms = float(sum(per_step))
        e2e_total = 0.0
        if e2e:
            step_e2e(c)
            barrier()
            e2e_total = float(sum(step_e2e(c) for _ in range(steps)))
            barrier()
        all_samples = c["out_samples"]
        if world > 1:
            t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, e2e_total = float(t[0]), float(t[1])
            n = torch.tensor([c["out_samples"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
            all_samples = int(n[0])
        return dict(value=all_samples * steps / (total_ms / 1e3), e2e_value=(all_samples * steps / (e2e_total / 1e3)) if e2e else None,
                    total_ms=total_ms, e2e_total=e2e_total, loop_s=float(np.mean(loop_ms)) / 1e3, last_t=last_t, launches=launches,
                    all_samples=all_samples)

    sampler = ClockSampler(local_rank)
    ctx = setup(wl, args.precision)
    sampler.start()
    m = measure(ctx, args.warmup, args.steps)
    clocks = sampler.finish()
    model = ctx["model"]
    mode, bits, seconds, batched, target, overlap = ctx["plan"]
    last_t = m["last_t"]
    floor = floor_us = None
    if rank == 0:
        floor = model.barrier_floor(20000)
        floor["cluster_us"] = model.cluster_floor(16, 20000)

    def loop_block(mres, precision):
        lt = mres["last_t"]
        kern = lt.get("loop_kernel", "?")
        n_exch = 8 if kern == "wrnn_loop_rr_kernel" else (6 if ctx["plan"][0] == "RAW" else 4) if kern == "wrnn_loop_rs_kernel" else 4 if kern == "wrnn_loop_gn_kernel" else (5 if (lt.get("precision") == _native.PREC_SPARSE_F32 or ctx["plan"][0] == "MOL") else 6)
        us = mres["loop_s"] * 1e6 / (lt["n_steps"] * max(1, lt["n_launches"]))
        fl = {"f32": floor["ll_us"], "f16": floor["counter_us"], "sparse": floor["cluster_us"]}[precision]
        extra = {}
        if kern == "wrnn_loop_rs_kernel":       # which conditioning form ran (engine default: inline; WRNN_RS_INLINE=0: record ring)
            extra["conditioning"] = ("records from expander CTAs through an L2-resident ring" if os.environ.get("WRNN_RS_INLINE") == "0" else
                                     "inline: per-frame rows + the mel share as a K = 80 slab of the on-path MMA (no per-sample records)")
        return {"kernel": kern, **extra, "us_per_step": us, "folds": lt["n_folds"], "loop_steps": lt["n_steps"], "exchanges_per_step": n_exch,
                "exchange_floor_us": floor["ll_us"], "counter_barrier_floor_us": floor["counter_us"], "cluster16_exchange_floor_us": floor["cluster_us"],
                "step_over_floor": us / (n_exch * fl), "step_over_flag_in_data_floor": us / (n_exch * floor["ll_us"]),
                "floor_used": {"f32": "flag-in-data exchange through L2", "f16": "fence+atomic counter barrier through L2",
                               "sparse": "DSMEM stores + cluster barrier (16 CTAs)"}[precision]}

    # ---- extra fold plans of config 3 (single GPU view; rank 0 prints them), in-run parity check, sharded runs ----------------
    plans, parity, sharded = {}, None, {}
    if not args.no_extras and wl == "cfg3ref":
        for other in ("cfg3a", "cfg3"):
            c2 = setup(other, "f16")
            m2 = measure(c2, 2, 2)
            if rank == 0:
                lb = loop_block(m2, "f16")
                plans[other] = {"workload": workload_config(other)["workload"], "value": m2["value"], "x_realtime": m2["value"] / 16000.0,
                                "e2e_x_realtime": m2["e2e_value"] / 16000.0, "n_gpus": world, "loop": lb}
            del c2
        ctx = setup(wl, args.precision)          # (the facade is a singleton: load the headline model again)
        model = ctx["model"]
    if not args.no_extras and rank == 0 and args.precision == "f16" and batched:
        # the benchmarked precision against the repo's fp32 parity-mode loop on THIS workload: the fp16 loop teacher-forced on the
        # fp32 loop's samples, first 48 steps of every fold (the oracle-anchored figures are tests/test_gpu_rs.py, test_gpu_tc.py)
        mel_n = ctx["mels_raw"][0] / np.float32(4.0)
        S_full = target + 2 * overlap
        a = model.generate_debug(mel_n, True, target, overlap, want_logits=True, seed=5, max_steps=48, precision=_native.PREC_F32)
        forced = np.zeros((a["samples"].shape[0], S_full), np.float32)
        forced[:, :48] = a["samples"]
        b = model.generate_debug(mel_n, True, target, overlap, forced=forced, want_logits=True, seed=5, max_steps=48, precision=_native.PREC_F16)
        rel = float(np.abs(a["logits"] - b["logits"]).max() / np.abs(a["logits"]).max())
        if mode == "MOL":
            agree = float((np.abs(a["samples"] - b["samples"]) < 1e-3).mean())
        else:
            agree = float((a["samples"] == b["samples"]).mean())
        parity = {"against": "fp32 parity-mode loop of this engine, teacher-forced, 48 steps x %d folds of this workload" % a["samples"].shape[0],
                  "logits_rel_err": rel, "draw_agreement": agree, "gates": {"logits_rel_err": 1e-3, "draw_agreement": 0.999}}
    if not args.no_extras and wl == "cfg3ref":
        # (1) strong scaling of the multi-utterance workload (BASELINE config 5: 256 utterances sharded by utterance over the ranks)
        c5 = setup("cfg5", "f16")
        m5 = measure(c5, 1, 2)
        if rank == 0:
            sharded["cfg5_strong"] = {"workload": workload_config("cfg5")["workload"], "scaling": "strong", "n_gpus": world,
                                      "value": m5["value"], "x_realtime": m5["value"] / 16000.0, "e2e_x_realtime": m5["e2e_value"] / 16000.0,
                                      "per_gpu": "256 utterances sharded by utterance index over the ranks, no collective on the data path",
                                      "loop": loop_block(m5, "f16")}
        del c5
        barrier()
        if world > 1:
            dist.barrier(group=cpu_group)        # (an NCCL barrier would park a spinning kernel on the GPUs rank 0 is about to use)
        # (2) ONE utterance, its folds split into contiguous ranges over all GPUs of the box, driven from one process (rank 0)
        #     through the public infer_waveform: host threads, host gather of the samples, post chain on GPU 0 -- all timed
        if world > 1 and rank == 0:
            try:
                hp = ctx["hp"]
                inference.load_state(ctx["sd"], devices=list(range(world)), override_hp_fatchord=hp)
                mel_h = ctx["mels_host"][0].numpy()
                keep_min, inference.SHARD_MIN_FOLDS = inference.SHARD_MIN_FOLDS, 0       # force the split (the facade itself keeps 213 folds on one GPU)
                inference.infer_waveform(mel_h, target=target, overlap=overlap)
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    w = inference.infer_waveform(mel_h, target=target, overlap=overlap)
                    ts.append(time.perf_counter() - t0)
                sharded["cfg3ref_fold_sharded"] = {"n_gpus": world, "x_realtime": (w.shape[0] / float(np.mean(ts))) / 16000.0,
                                                   "ms_per_call": float(np.mean(ts)) * 1e3,
                                                   "what": "one 60 s utterance, 213 folds split into %d contiguous ranges, one engine + host thread per "
                                                           "GPU, host gather + crossfade on GPU 0 inside the timed region (inference._infer_sharded, "
                                                           "forced: at <= 256 folds a step is a latency chain that does not shorten with fewer "
                                                           "folds, so the facade keeps such an utterance on one GPU)" % world}
                inference.SHARD_MIN_FOLDS = keep_min
            except Exception as e:
                sharded["cfg3ref_fold_sharded"] = {"error": str(e)[:300]}
        if world > 1:
            dist.barrier(group=cpu_group)
        barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value, e2e_value = m["value"], m["e2e_value"]
    pk, pk_kind = peaks()
    C = model.n_classes
    F, S = last_t["n_folds"], last_t["n_steps"]
    loop_s = m["loop_s"]
    kern = last_t.get("loop_kernel", "?")
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s:%s" % (wl, kern))
    except Exception:
        pass
    if batched and args.precision == "f16":
        # tensor bound: algorithmic FLOPs of the REFERENCE's step over the loop kernel's time
        flops = 2.0 * topo_macs(wl, C) * F * S
        achieved = flops / loop_s / 1e12
        peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")))
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": pk_kind + " bf16_tflops_sustained", "algorithmic_flops_per_launch": flops}
    else:
        # weight-streaming bound of the fp32 loops (CUDA-core FMAs, no tensor pipe): every step touches the fp32 loop
        # weights once, whatever the number of folds; they live on-chip here
        wbytes = 4.0 * topo_macs(wl, C)
        if wl in PRUNED:
            wbytes = 4.0 * 0.41e6 * 1.25            # ~0.41 M MAC per fold-step at 90 % sparsity: value + one index byte per 1x4 group
        achieved = wbytes * S * max(1, last_t["n_launches"]) * max(1, F if wl in PRUNED else 1) / loop_s / 1e9
        peak = float(pk.get("hbm_gbs"))
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": pk_kind + " hbm_gbs", "algorithmic_bytes_per_step": wbytes,
                "note": "weight-streaming bound of SURVEY.md 8(d): algorithmic weight bytes per step x steps / loop time; the weights are "
                        "resident in shared memory, so this can exceed nothing -- the exchange floor below is the figure of merit"}
    roof.update({"traffic": (traffic or {}).get("bytes"), "traffic_source": (traffic or {}).get("source"), "kernel": kern, "kernel_ms": loop_s * 1e3})
    line = {
        "metric": "vocoder_output_samples_per_sec", "value": value, "unit": "samples/s", "x_realtime": value / 16000.0,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["total_ms"] / args.steps,
        "higher_is_better": True, "scaling": "strong" if wl in MULTI else "weak", "vs_baseline": None,
        "dtype": "f16" if args.precision == "f16" else "f32", "data": "synthetic",
        "precision_note": ("fp32 weights/FMA/state (parity mode)" if args.precision == "f32" else
                           "fp32 block-sparse (1x4 groups) cluster-local loop" if args.precision == "sparse" else
                           "fp16 weights+activations on tcgen05, fp32 accumulate, fp32 recurrent state and conditioning; parity measured in "
                           "this run: see parity_check (oracle-anchored gates: tests/test_gpu_rs.py, tests/test_gpu_tc.py)"),
        "config": dict(workload_config(wl), folds=F, loop_steps=S, per_gpu=("256 utterances sharded by utterance" if wl in MULTI else "one utterance per GPU, independent"),
                       pruned=("~90% zero 1x4 groups (pruner.py rule), block-sparse cluster loop (loop_sparse.cu)" if wl in PRUNED else None)),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "x_realtime": e2e_value / 16000.0,
                "h2d_bytes_per_step": int(sum(mm.nbytes for mm in ctx["mels_raw"])), "d2h_bytes_per_step": int(ctx["out_samples"] * 8),
                "ms_per_step": m["e2e_total"] / args.steps},
        "gpu_launches": int(m["launches"]),
        "roofline": roof,
        "loop": loop_block(m, args.precision),
        "phases_ms": {k: last_t[k] for k in ("ms_h2d", "ms_cond", "ms_loop", "ms_post", "ms_d2h")},
    }
    if plans:
        line["plans"] = plans
    if parity:
        line["parity_check"] = parity
    if sharded:
        line["sharded"] = sharded
    if not args.no_cpu_baseline:
        # the CPU arm runs in its own process with the GPUs hidden (the reference moves to a GPU whenever it sees one)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", wl, "--steps", "1", "--warmup", "0",
                                "--cpu-seconds", str(args.cpu_seconds)], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1"))
            line["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:
            line["cpu_baseline"] = {"unavailable": str(e)[:300]}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
