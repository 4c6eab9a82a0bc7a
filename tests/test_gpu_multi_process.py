"""Real multi-process, multi-GPU runs (one process per GPU, NCCL for the plumbing, cudaIpc for the gradient windows).
Skipped on a single-GPU box; on >= 2 GPUs it checks SURVEY Appendix B item 4: the data-parallel large-batch step with the
peer-memory all-reduce keeps the replicas bit-identical and agrees with the single-rank step of the same global batch,
and a population sharded over two processes equals the unsharded one bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["DQN_REPO_ROOT"])
import dqn_b200
from oracle import dqn_oracle as O
from oracle.replay_oracle import synthetic_transitions

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
out = os.environ["DQN_TEST_OUT"]
HID, B, D, A, N = (256, 256), 512, 8, 4, 3000
rng = np.random.default_rng(0)
params = O.init_params(rng, D, A, hidden=HID, bias_std=0.05)
data = synthetic_transitions(rng, 2500, D, A, done_p=0.2)

def run(collective):
    tr = dqn_b200.LargeBatchTrainer(D, A, HID, B, N, 0.99, dqn_b200.adamw(1e-3), rank=rank, world_size=world, seed=7, device=rank,
                                    collective="p2p" if collective == "p2p_seq" else collective)
    tr.set_params(params, 0); tr.set_params(params, 1)
    tr.store(*data)
    for _ in range(4):
        if collective == "p2p_seq":          # the exchange as ONE whole-vector kernel behind backward (no overlap)
            tr.forward_backward(); tr.all_reduce(); tr.apply()
        else:                                # p2p: dqn_lb_train_step, W2 gradient exchanged on a second stream under dh1 / dW1
            tr.step()
    flat = np.concatenate([np.ravel(tr.get_params()[m][k]) for m in O.MODULES for k in ("w", "b")])
    return flat, tr.loss()

for coll in ("p2p", "p2p_seq", "nccl"):
    flat, loss = run(coll)
    t = torch.from_numpy(flat).cuda()
    every = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(every, t)
    if rank == 0:
        np.savez(os.path.join(out, f"dp_{coll}.npz"), params=np.stack([e.cpu().numpy() for e in every]), loss=loss)

# population: 8 agents over `world` processes
pop = dqn_b200.Population(8, D, A, 500, dqn_b200.adam(1e-3), rank=rank, world_size=world, seed=3, device=rank)
for i in range(pop.n_local):
    g = pop.global_id(i)
    pop.store(i, *synthetic_transitions(np.random.default_rng([5, g]), 300, D, A))
pop.train_steps(5)
mine = np.stack([pop.params_flat(i) for i in range(pop.n_local)])
t = torch.from_numpy(mine).cuda()
every = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(every, t)
if rank == 0:
    np.savez(os.path.join(out, "pop.npz"), params=np.concatenate([e.cpu().numpy() for e in every]))
dist.destroy_process_group()
'''


def test_two_processes_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import dqn_b200
    from oracle import dqn_oracle as O
    from oracle.replay_oracle import synthetic_transitions
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DQN_REPO_ROOT=ROOT, DQN_TEST_OUT=str(tmp_path))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=500)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    # ---- large-batch DP: replicas bit-identical; p2p ~ nccl ~ single rank ----
    p2p, nccl = np.load(tmp_path / "dp_p2p.npz"), np.load(tmp_path / "dp_nccl.npz")
    assert np.array_equal(p2p["params"][0], p2p["params"][1]) and np.array_equal(nccl["params"][0], nccl["params"][1])
    seq = np.load(tmp_path / "dp_p2p_seq.npz")
    assert np.array_equal(seq["params"], p2p["params"])          # overlapped exchange == whole-vector exchange, bit for bit
    HID, B, D, A, N = (256, 256), 512, 8, 4, 3000
    rng = np.random.default_rng(0)
    params = O.init_params(rng, D, A, hidden=HID, bias_std=0.05)
    data = synthetic_transitions(rng, 2500, D, A, done_p=0.2)
    one = dqn_b200.LargeBatchTrainer(D, A, HID, B, N, 0.99, dqn_b200.adamw(1e-3), seed=7)
    one.set_params(params, 0); one.set_params(params, 1)
    one.store(*data)
    for _ in range(4):
        one.step()
    ref = np.concatenate([np.ravel(one.get_params()[m][k]) for m in O.MODULES for k in ("w", "b")])
    # four Adam steps at lr 1e-3 from the same start: shard-sum vs single-rank gradients differ by fp32 summation order only
    for got in (p2p["params"][0], nccl["params"][0]):
        assert np.abs(got - ref).max() <= 5e-5 * max(1.0, np.abs(ref).max())
    assert abs(float(p2p["loss"]) - one.loss()) <= 1e-5 * abs(one.loss())
    # ---- population: sharded == unsharded, bit for bit ----
    pop = dqn_b200.Population(8, D, A, 500, dqn_b200.adam(1e-3), rank=0, world_size=1, seed=3)
    for i in range(8):
        pop.store(i, *synthetic_transitions(np.random.default_rng([5, i]), 300, D, A))
    pop.train_steps(5)
    whole = np.stack([pop.params_flat(i) for i in range(8)])
    assert np.array_equal(np.load(tmp_path / "pop.npz")["params"], whole)
