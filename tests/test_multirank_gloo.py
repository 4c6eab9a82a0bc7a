"""world_size-2 gloo run of the N>1 plumbing bench.py uses: contiguous agent shards with no data-path
collective, a barrier, and a MAX reduction of the per-rank time."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import dqn_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 37
    b, e = dqn_b200.shard_range(n, rank, world)
    hps = dqn_b200.sweep_hparams(n)[b:e]
    # per-shard checksum of the hyper-parameters; gathered only for the test (the data path has no collective)
    local = torch.tensor([sum(h["batch_size"] for h in hps), e - b], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, local)
    dist.barrier()
    t = torch.tensor([1.0 + rank], dtype=torch.float64)          # stand-in for the device time of this rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(([g.tolist() for g in gathered], float(t)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_plumbing():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, tmax = out.get()
    import dqn_b200
    full = dqn_b200.sweep_hparams(37)
    assert sum(g[1] for g in gathered) == 37
    assert sum(g[0] for g in gathered) == sum(h["batch_size"] for h in full)
    assert tmax == 2.0                                             # max over ranks, not rank 0's own time
