"""Cost of one agent-step of the population kernel as a function of the batch size (592 agents = 4 waves of 148 SMs, every agent
the same batch; K = 128 steps per launch).  usage: python profiles/pop_batch_cost.py [cta|cta_tc] [batch ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
kern = sys.argv[1] if len(sys.argv) > 1 else "cta_tc"
os.environ["DQN_B200_STEP_KERNEL"] = kern
import torch, dqn_b200
from oracle.replay_oracle import synthetic_transitions
n, K = 592, 128
rng = np.random.default_rng(0)
data = synthetic_transitions(rng, 4000, 8, 4, done_p=0.05)
batches = [int(x) for x in sys.argv[2:]] or [38, 54, 64, 65, 70, 80, 96, 128]
for B in batches:
    e = dqn_b200.DqnEngine(8, 4, 4000, B, 0.99, dqn_b200.adam(1e-4), n_agents=n, seed=1)
    theta = dqn_b200.Model(4).init(np.random.default_rng(3), np.zeros((1, 8), np.float32))
    flat = dqn_b200.pkg.specs.flatten_tree(theta, 8, 4) if hasattr(dqn_b200.pkg, "specs") else None
    for i in range(n):
        e.store(*data, agent=i)
    e.train_steps(K, agent_begin=0, agent_end=n); e.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3): e.train_steps(K, agent_begin=0, agent_end=n)
    e.synchronize(); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print(f"{kern} batch {B:3d}: {ms:7.3f} ms per launch, {ms * 1e3 / K / 4:6.2f} us per agent-step per SM, {n * K / ms / 1e3:6.2f} M agent-steps/s")
    e.close()
