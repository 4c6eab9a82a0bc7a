"""Host-side cost of the e2e loop of bench.py (adds / _step / loss read), per step, and the loop with the loss read one step
behind (two commands in flight).  usage: python profiles/host_breakdown.py [--no-session]"""
import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import bench, dqn_b200, torch
agent, data = bench.build_agent(dqn_b200, 0, 0, session='--no-session' not in sys.argv)
eng = agent._engine; rb = agent._replay_buffer
s,a,r,s2,d = [x[:80000] for x in data]
a_py=[int(x) for x in a]; r_py=[float(x) for x in r]; d_py=[bool(x) for x in d]
def loop(n, off, lag):
    ta=tb=tc=0.0
    for i in range(n):
        t0=time.perf_counter()
        for j in range(4):
            k=off+i*4+j
            rb.add(s[k],a_py[k],r_py[k],s2[k],d_py[k])
        t1=time.perf_counter()
        agent._step()
        t2=time.perf_counter()
        if i: rb.last_loss(lag)
        t3=time.perf_counter()
        ta+=t1-t0; tb+=t2-t1; tc+=t3-t2
    rb.last_loss()
    return ta/n*1e6,tb/n*1e6,tc/n*1e6
def plain(n, off, lag):
    t0=time.perf_counter()
    for i in range(n):
        for j in range(4):
            k=off+i*4+j
            rb.add(s[k],a_py[k],r_py[k],s2[k],d_py[k])
        agent._step()
        if i: rb.last_loss(lag)
    rb.last_loss()
    return (time.perf_counter()-t0)/n*1e6
loop(200,0,0)
for lag in (0, 1):
    print("lag %d: adds %.2f us  _step %.2f us  loss wait %.2f us" % ((lag,) + loop(4000,800+lag*16000,lag)))
    print("lag %d: untimed loop %.2f us per step" % (lag, plain(4000, 40000+lag*16000, lag)))
# the adds alone (no step): pure host staging
t0=time.perf_counter()
for i in range(4000):
    for j in range(4):
        k=i*4+j
        rb.add(s[k],a_py[k],r_py[k],s2[k],d_py[k])
    rb._pending = 0
print("adds only %.2f us per step" % ((time.perf_counter()-t0)/4000*1e6))
