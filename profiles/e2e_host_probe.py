"""Is the e2e loop (4 x ReplayBuffer.add + Agent._step + lagged loss read per step, Agent(session=True)) paced by the host?
Runs the loop of bench.py three ways: (a) as bench.py does, (b) with the state rows pre-split into a list (no numpy row-view
creation inside the loop), (c) calling the C staging helper directly (no Python frames for add / _step / last_loss).
usage: python profiles/e2e_host_probe.py [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench, dqn_b200
from dqn_b200 import pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
agent, data = bench.build_agent(dqn_b200, 0, seed=0, session=True)
rb, eng = agent._replay_buffer, agent._engine
TF = bench.TRAIN_FREQUENCY
s, a, r, s2, d = [x[:TF * (4 * n + 64)] for x in data]
a_py, r_py, d_py = a.tolist(), r.tolist(), d.tolist()
s_rows, s2_rows = list(s), list(s2)
hs = pkg.replay._hoststage

def loop_a(n, off):
    for i in range(n):
        for j in range(TF):
            k = off + i * TF + j
            rb.add(s[k], a_py[k], r_py[k], s2[k], d_py[k])
        agent._step()
        if i: rb.last_loss(1)
    return rb.last_loss()

def loop_b(n, off):
    for i in range(n):
        for j in range(TF):
            k = off + i * TF + j
            rb.add(s_rows[k], a_py[k], r_py[k], s2_rows[k], d_py[k])
        agent._step()
        if i: rb.last_loss(1)
    return rb.last_loss()

def loop_c(n, off):
    st, put, step, loss = rb._stage, hs.put, hs.step, hs.loss
    for i in range(n):
        k = off + i * TF
        put(st, 0, s_rows[k], a_py[k], r_py[k], s2_rows[k], d_py[k])
        put(st, 1, s_rows[k + 1], a_py[k + 1], r_py[k + 1], s2_rows[k + 1], d_py[k + 1])
        put(st, 2, s_rows[k + 2], a_py[k + 2], r_py[k + 2], s2_rows[k + 2], d_py[k + 2])
        put(st, 3, s_rows[k + 3], a_py[k + 3], r_py[k + 3], s2_rows[k + 3], d_py[k + 3])
        step(st, 4)
        if i: loss(st, 1)
    rb._counter += 4 * n
    return loss(st, 0)

off = 0
for name, fn in (("a: bench.py's loop", loop_a), ("b: rows pre-split", loop_b), ("c: C helper called directly", loop_c), ("a again", loop_a)):
    fn(8, off); off += 8 * TF
    eng.synchronize()
    t0 = time.perf_counter()
    fn(n, off); off += n * TF
    eng.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:32s} {dt / n * 1e6:6.2f} us per step  {n / dt / 1e3:7.1f} k steps/s")
