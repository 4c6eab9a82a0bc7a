"""Generate ``tests/golden/*.npz`` from the REAL reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only) and numba

The reference cannot travel to the GPU box, so its outputs are frozen here as small fixtures:

* ``ref_checkpoint.npz``   -- ``Test/lunar_lander/{params,opt_state}.pickle`` converted leaf by
  leaf to numpy (tree names, [in,out] layouts, dtypes, theta_0, Adam ``(count, mu, nu)``).
* ``replay_ref_stream.npz`` -- a seeded stream of transitions pushed through the reference's own
  ``ReplayBuffer.add`` (``General/Base/replay_buffer.py:58-65``), including wrap-around, with the
  five arrays / ``size`` / ``_counter`` captured at several points.
* ``replay_ref_sample.npz`` -- the reference's numba ``sample_batch`` (``replay_buffer.py:68-85``)
  run with numba's MT19937 seeded from inside an ``@njit`` helper; the drawn indices are
  recovered from a tag planted in ``states[:, 0]``.

Nothing in ``tests -m gpu``, ``smoke()`` or ``bench.py`` reads /root/reference; they read these files.
"""
import os
import pickle
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _reconstruct_device_array(fun, args, arr_state, aval_state):
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return arr


class _RefUnpickler(pickle.Unpickler):
    """jax/optax are absent: map DeviceArray -> numpy and optax NamedTuples -> plain tuples."""

    def find_class(self, module, name):
        if module.startswith("jax") and name == "reconstruct_device_array":
            return _reconstruct_device_array
        if module.startswith("optax"):
            return type(name, (tuple,), {"__new__": lambda cls, *a: tuple.__new__(cls, a)})
        return super().find_class(module, name)


def checkpoint():
    d = os.path.join(REF, "Test", "lunar_lander")
    with open(os.path.join(d, "params.pickle"), "rb") as f:
        params = _RefUnpickler(f).load()
    with open(os.path.join(d, "opt_state.pickle"), "rb") as f:
        opt_state = _RefUnpickler(f).load()
    out = {"module_order": np.array(list(params.keys()))}
    for mod, leaves in params.items():
        for k, v in leaves.items():
            out[f"params|{mod}|{k}"] = np.asarray(v)
    adam = opt_state[0]
    assert type(adam).__name__ == "ScaleByAdamState" and len(opt_state) == 3
    out["opt_chain"] = np.array([type(x).__name__ for x in opt_state])
    out["count"] = np.asarray(adam[0])
    for nm, tree in (("mu", adam[1]), ("nu", adam[2])):
        for mod, leaves in tree.items():
            for k, v in leaves.items():
                out[f"{nm}|{mod}|{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "ref_checkpoint.npz"), **out)
    print("ref_checkpoint.npz:", len(out), "entries")


def replay():
    sys.path.insert(0, REF)
    import numba
    from General.Base.replay_buffer import ReplayBuffer, sample_batch   # the reference's own code

    # --- add() stream with wrap-around -----------------------------------------------------
    N, D, total = 37, 9, 100
    rng = np.random.default_rng(20221018)
    s = rng.standard_normal((total, D)).astype(np.float32)
    a = rng.integers(0, 4, total)
    r = rng.standard_normal(total) * 2.0                 # python floats (f64) -> stored as f32
    s2 = rng.standard_normal((total, D)).astype(np.float32)
    d = rng.random(total) < 0.2
    buf = ReplayBuffer(N, (N, D), (N,))
    out = dict(N=N, D=D, in_states=s, in_actions=a, in_rewards=r, in_observations=s2, in_dones=d)
    marks = [1, 20, 37, 38, 64, 100]
    for i in range(total):
        buf.add(s[i], int(a[i]), float(r[i]), s2[i], bool(d[i]))
        if i + 1 in marks:
            k = f"after{i + 1}"
            out[f"{k}_states"] = buf.states.copy()
            out[f"{k}_actions"] = buf.actions.copy()
            out[f"{k}_rewards"] = buf.rewards.copy()
            out[f"{k}_observations"] = buf.observations.copy()
            out[f"{k}_dones"] = buf.dones.copy()
            out[f"{k}_size"] = buf.size
            out[f"{k}_counter"] = buf._counter
    out["marks"] = np.array(marks)
    np.savez_compressed(os.path.join(OUT, "replay_ref_stream.npz"), **out)
    print("replay_ref_stream.npz ok; dtypes:", buf.states.dtype, buf.actions.dtype,
          buf.rewards.dtype, buf.observations.dtype, buf.dones.dtype)

    # --- sample_batch() with numba's generator seeded ---------------------------------------
    @numba.njit
    def seed_numba(x):
        np.random.seed(x)

    N, D = 500, 9
    rng = np.random.default_rng(7)
    buf = ReplayBuffer(N, (N, D), (N,))
    filled = 321                                         # partially filled: sample from the prefix only
    for i in range(filled):
        st = rng.standard_normal(D).astype(np.float32)
        st[0] = i                                        # tag: recover the drawn index
        buf.add(st, int(rng.integers(0, 4)), float(rng.standard_normal()),
                rng.standard_normal(D).astype(np.float32), bool(rng.random() < 0.2))
    out = dict(N=N, D=D, filled=filled, states=buf.states.copy(), actions=buf.actions.copy(),
               rewards=buf.rewards.copy(), observations=buf.observations.copy(),
               dones=buf.dones.copy())
    for case, (seed, B) in enumerate([(123, 64), (5, 38), (99, 70), (1, 1)]):
        seed_numba(seed)
        bs, ba, br, bo, bd = sample_batch(buf.size, buf.states, buf.actions, buf.rewards,
                                          buf.observations, buf.dones, B)
        idx = bs[:, 0].astype(np.int64)
        assert idx.min() >= 0 and idx.max() < filled
        out[f"case{case}_B"] = B
        out[f"case{case}_idx"] = idx
        out[f"case{case}_states"] = bs
        out[f"case{case}_actions"] = ba
        out[f"case{case}_rewards"] = br
        out[f"case{case}_observations"] = bo
        out[f"case{case}_dones"] = bd
    out["n_cases"] = 4
    np.savez_compressed(os.path.join(OUT, "replay_ref_sample.npz"), **out)
    print("replay_ref_sample.npz ok; batch dtypes:", bs.dtype, ba.dtype, br.dtype, bo.dtype, bd.dtype)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    checkpoint()
    replay()
