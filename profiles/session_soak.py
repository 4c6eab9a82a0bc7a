import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import dqn_b200
from oracle import dqn_oracle as O
from oracle.replay_oracle import synthetic_transitions
_lib = dqn_b200.pkg._lib
theta = O.init_params(np.random.default_rng(1), 8, 4, bias_std=0.05)
engs = []
for session in (False, True):
    e = dqn_b200.DqnEngine(8, 4, 5000, 64, 0.99, dqn_b200.adamw(2e-4), seed=3, step_kernel="cluster", session=session)
    e.set_params(theta, 0, 0); e.set_params(theta, 0, 1); engs.append(e)
rng = np.random.default_rng(0)
n = 100000
s, a, r, s2, d = synthetic_transitions(rng, 4 * n, 8, 4, done_p=0.05)
d8 = np.ascontiguousarray(d, dtype=np.bool_)
loss = np.zeros(1, np.float32)
losses = [[], []]
for which, e in enumerate(engs):
    t0 = time.perf_counter()
    for i in range(n):
        k = 4 * i
        if e is engs[0]:
            _lib.check(e.lib.dqn_store_train_step(e.h, 0, 4, _lib.ptr(s[k:k+4]), _lib.ptr(a[k:k+4]), _lib.ptr(r[k:k+4]), _lib.ptr(s2[k:k+4]), _lib.ptr(d8[k:k+4]), 1, _lib.ptr(loss)))
            if i % 97 == 0: losses[0].append(float(loss[0]))
        else:      # the pipelined form: publish step i, then read step i - 1's loss (two commands in flight)
            _lib.check(e.lib.dqn_store_train_step(e.h, 0, 4, _lib.ptr(s[k:k+4]), _lib.ptr(a[k:k+4]), _lib.ptr(r[k:k+4]), _lib.ptr(s2[k:k+4]), _lib.ptr(d8[k:k+4]), 1, None))
            if i and (i - 1) % 97 == 0: losses[1].append(e.lagged_loss(0, 1))
            elif i: e.lagged_loss(0, 1)
        if i % 5 == 0:
            e.act(s[k])
        if i % 1000 == 999:
            e.sync_target()
    e.synchronize()
    print("session" if e is engs[1] else "launch ", "%.1f us per step" % ((time.perf_counter() - t0) / n * 1e6), "loss", loss[0])
assert losses[0][:len(losses[1])] == losses[1], "lagged losses differ"
assert np.array_equal(engs[0].get_params_flat(), engs[1].get_params_flat())
assert np.array_equal(engs[0].get_params_flat(0, 1), engs[1].get_params_flat(0, 1))
assert engs[0].get_counters() == engs[1].get_counters()
print("100k steps: session == launch-per-call bit for bit")
