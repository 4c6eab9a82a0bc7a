#!/usr/bin/env python
"""Where a K=1 launch of the cluster train-step kernel spends its cycles (rank 0, thread 0, SM clock).
Needs the profiling variant: make -C deep-q-learning_b200/csrc variants/libdqn_phaseclk.so
usage: DQN_B200_LIB=deep-q-learning_b200/csrc/variants/libdqn_phaseclk.so python profiles/phase_clocks.py [K]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, dqn_b200, torch
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
agent, data = bench.build_agent(dqn_b200, 0, 0)
lib = agent._engine.lib
names = ["entry", "prologue issued", "cluster.sync+prefetch", "first gather landed", "tiles done (last step)",
         "cluster.sync #1", "Adam + sync #2 + G zero", "write-back issued"]
acc = np.zeros(8)
n = 200
for i in range(n + 20):
    agent._steps(K)
    torch.cuda.synchronize()
    out = (C.c_longlong * 16)()
    lib.dqn_debug_phase_clocks(out)
    c = np.array(out[:8], dtype=np.float64)
    if i >= 20:
        acc += c - c[0]
acc /= n
prev = 0.0
for nm, v in zip(names, acc):
    print(f"{nm:28s} {v:9.0f} cyc  (+{v - prev:7.0f})")
    prev = v
