"""N>1 host logic on CPU (gloo, world_size 2): the path shards with no data-path collective, so what needs
checking is the partition itself -- every utterance / fold is owned by exactly one rank, the Philox counters are
global (independent of the sharding) and the timing reduction bench.py uses (MAX over ranks, SUM of samples)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import philox, wavernn_oracle as orc
    # utterance sharding of cfg5
    mels, idx = bench.workload_mels("cfg5", rank, world)
    mine = torch.zeros(256, dtype=torch.int64)
    mine[idx] = 1
    frames = torch.tensor([sum(m.shape[1] for m in mels)], dtype=torch.float64)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    fsum = frames.clone()
    dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
    fmax = frames.clone()
    dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
    # fold-range sharding of one utterance (inference._infer_sharded): contiguous, disjoint, complete
    F, _ = orc.fold_plan(960000, 3000, 1500)
    bounds = [F * i // world for i in range(world + 1)]
    cover = torch.zeros(F, dtype=torch.int64)
    cover[bounds[rank]:bounds[rank + 1]] = 1
    dist.all_reduce(cover, op=dist.ReduceOp.SUM)
    # noise of my fold range == the same rows of the global noise (counters are (step, fold, utterance))
    u_mine = philox.raw_uniforms(7, 5, bounds[rank + 1] - bounds[rank], utt=3, fold0=bounds[rank])
    u_all = philox.raw_uniforms(7, 5, F, utt=3)
    same = bool(np.array_equal(u_mine, u_all[:, bounds[rank]:bounds[rank + 1]]))
    # bench timing reduction
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(dict(owners=mine.tolist(), imbalance=float(fmax / (fsum / world)), cover=cover.tolist(), same=same,
                     tmax=float(t)))
    else:
        out.put(dict(same=same))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharding_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r["same"] for r in res)
    full = [r for r in res if "owners" in r][0]
    assert full["owners"] == [1] * 256                    # every utterance owned exactly once
    assert full["cover"] == [1] * len(full["cover"])      # every fold owned exactly once
    assert full["imbalance"] < 1.1
    assert full["tmax"] == 11.0
