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


def _dp_worker(rank, world, port, out):
    """The data-parallel large-batch step on the CPU oracle: every rank differentiates its contiguous slice of the global
    minibatch with the loss scaled by 1/B_global, the P + 1 floats (gradients + loss share) are summed with ONE
    all-reduce, every rank applies the identical Adam update (what LargeBatchTrainer.step does with its kernels)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dqn_oracle as O
    from oracle.philox import sample_indices
    from oracle.replay_oracle import synthetic_transitions
    D, A, Bg, N = 8, 4, 64, 500
    rng = np.random.default_rng(0)                                  # same replay replica and weights on every rank
    params = O.init_params(rng, D, A, bias_std=0.05)
    target = O.tree_map(lambda x: (x + 0.02 * rng.standard_normal(x.shape)).astype(np.float32), params)
    data = synthetic_transitions(rng, N, D, A, done_p=0.2)
    opt, opt_state = O.OptSpec("adamw", 1e-3), O.init_opt_state(params)
    bl = Bg // world
    for step in range(3):
        idx = sample_indices(7, 0, step, Bg, N)[rank * bl:(rank + 1) * bl]        # this rank's slice of the global draw
        s, a, r, s2, d = O.preprocessing(*[x[idx] for x in data])
        tgt = O.compute_q_targets(params, target, s, a, r, s2, d, 0.99)
        loss, grads = O.loss_and_grads(params, s, tgt)                             # mean over the LOCAL rows ...
        flat = np.concatenate([np.ravel(grads[m][k]) for m in O.MODULES for k in ("w", "b")] + [[loss]]).astype(np.float64)
        t = torch.from_numpy(flat * (bl / Bg))                                     # ... rescaled to 1 / B_global
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        red = t.numpy().astype(np.float32)
        g, o = O.tree_zeros_like(params), 0
        for m in O.MODULES:
            for k in ("w", "b"):
                n = g[m][k].size
                g[m][k] = red[o:o + n].reshape(g[m][k].shape)
                o += n
        params, opt_state = O.adam_update(params, g, opt_state, opt)
    flat_p = np.concatenate([np.ravel(params[m][k]) for m in O.MODULES for k in ("w", "b")])
    every = [torch.zeros(flat_p.size, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(every, torch.from_numpy(flat_p))
    if rank == 0:
        out.put(([e.numpy() for e in every], float(red[-1])))
    dist.destroy_process_group()


def test_two_rank_data_parallel_step_equals_single_rank_on_the_oracle():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    replicas, loss = out.get()
    assert np.array_equal(replicas[0], replicas[1])                 # every rank consumed the same reduced buffer
    sys.path.insert(0, ROOT)
    from oracle import dqn_oracle as O
    from oracle.agent_oracle import OracleAgent
    from oracle.replay_oracle import synthetic_transitions
    rng = np.random.default_rng(0)
    params = O.init_params(rng, 8, 4, bias_std=0.05)
    target = O.tree_map(lambda x: (x + 0.02 * rng.standard_normal(x.shape)).astype(np.float32), params)
    data = synthetic_transitions(rng, 500, 8, 4, done_p=0.2)
    one = OracleAgent(params, O.init_opt_state(params), O.OptSpec("adamw", 1e-3), 500, 8, 0.99, 64, seed=7)
    one.target_params = target
    one.replay.add_many(*data)
    for _ in range(3):
        last = one.step()
    ref = np.concatenate([np.ravel(one.params[m][k]) for m in O.MODULES for k in ("w", "b")])
    assert np.abs(replicas[0] - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())     # fp32 summation order only
    assert abs(loss - float(last["loss"])) <= 1e-5 * abs(float(last["loss"]))
