import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import bench, dqn_b200, torch
agent, data = bench.build_agent(dqn_b200, 0, 0, session='--no-session' not in sys.argv)
eng = agent._engine; rb = agent._replay_buffer
s,a,r,s2,d = [x[:40000] for x in data]
a_py=[int(x) for x in a]; r_py=[float(x) for x in r]; d_py=[bool(x) for x in d]
def loop(n, off):
    ta=tb=tc=0.0
    for i in range(n):
        t0=time.perf_counter()
        for j in range(4):
            k=off+i*4+j
            rb.add(s[k],a_py[k],r_py[k],s2[k],d_py[k])
        t1=time.perf_counter()
        agent._step()
        t2=time.perf_counter()
        eng.last_loss()
        t3=time.perf_counter()
        ta+=t1-t0; tb+=t2-t1; tc+=t3-t2
    return ta/n*1e6,tb/n*1e6,tc/n*1e6
loop(200,0)
print("adds %.2f us  _step %.2f us  last_loss wait %.2f us" % loop(4000,800))
