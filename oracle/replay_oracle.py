"""CPU restatement of the reference replay ring (oracle; test infra only).  PARITY PINNED.

Follows ``General/Base/replay_buffer.py``: constructor ``:20-32`` (five SoA arrays: ``states
f32[obs_shape]``, ``actions i64[ac_shape]``, ``rewards f32[N]``, ``observations f32[obs_shape]``,
``dones bool[N]``), ``add`` ``:58-65`` (slot ``counter % N``; ``counter += 1``;
``num_samples = min(counter, N)``) and ``sample_batch`` ``:68-85`` (``B`` uniform indices in
``[0, num_samples)`` with replacement, then five fancy-index gathers).

Checked against the real reference module (numba 0.65) by ``oracle/make_golden.py`` -> the vectors
in ``tests/golden/replay_*.npz``; ``tests/test_oracle_replay.py`` replays them.
"""
import numpy as np


class OracleReplay:
    def __init__(self, buffer_size, obs_shape, ac_shape):
        self.buffer_size = int(buffer_size)
        self.states = np.zeros(obs_shape, dtype=np.float32)
        self.actions = np.zeros(ac_shape, dtype=np.int64)
        self.rewards = np.zeros((buffer_size,), dtype=np.float32)
        self.observations = np.zeros(obs_shape, dtype=np.float32)
        self.dones = np.zeros((buffer_size,), dtype=np.bool_)
        self.counter = 0
        self.size = 0

    def add(self, state, action, reward, observation, done):
        pos = self.counter % self.buffer_size
        self.states[pos] = state
        self.actions[pos] = action
        self.rewards[pos] = reward
        self.observations[pos] = observation
        self.dones[pos] = done
        self.counter += 1
        self.size = min(self.counter, self.buffer_size)

    def add_many(self, states, actions, rewards, observations, dones):
        """n scalar ``add`` calls in order (what a vectorised store must be equivalent to)."""
        for i in range(len(actions)):
            self.add(states[i], actions[i], rewards[i], observations[i], dones[i])

    def arrays(self):
        return self.states, self.actions, self.rewards, self.observations, self.dones


def gather(indices, states, actions, rewards, observations, dones):
    """The five fancy-index gathers of ``sample_batch`` (``replay_buffer.py:78-84``)."""
    idx = np.asarray(indices, dtype=np.int64)
    return states[idx], actions[idx], rewards[idx], observations[idx], dones[idx]


def sample_batch(rng, num_samples, states, actions, rewards, observations, dones, batch_size):
    """``sample_batch`` with an explicit generator standing in for numba's private MT19937."""
    idx = rng.integers(0, num_samples, batch_size, dtype=np.int64)      # replay_buffer.py:77
    return gather(idx, states, actions, rewards, observations, dones)


def numba_sample_batch():
    """The reference's sampler on the reference's engine: a numba-njit function drawing with numba's own
    ``numpy.random.randint`` (its private MT19937) and gathering the five arrays by fancy index
    (``replay_buffer.py:68-85``).  Used by ``bench.py``'s CPU baseline on boxes where ``/root/reference`` does not
    exist, so that the baseline's replay half costs what the reference's costs (the NumPy/Philox ``sample_indices`` is ~6x
    slower and is only there for index parity with the kernels).  Compiled on first call; returns the jitted function."""
    import numba

    @numba.njit
    def sample(num_samples, states, actions, rewards, observations, dones, batch_size):
        idx = np.random.randint(0, num_samples, batch_size)
        return states[idx], actions[idx], rewards[idx], observations[idx], dones[idx]

    return sample


def reference_replay_module(path="/root/reference/General/Base/replay_buffer.py"):
    """The reference's own ``replay_buffer`` module (real ``ReplayBuffer.add`` + numba ``sample_batch``), imported
    unmodified from where it lies -- only possible in the build container; ``None`` elsewhere."""
    import importlib.util
    import os
    if not os.path.exists(path):
        return None
    try:
        spec = importlib.util.spec_from_file_location("_reference_replay_buffer", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


def synthetic_transitions(rng, n, obs_dim, num_actions=4, done_p=0.01):
    """Synthetic transitions of SURVEY section 8(d): s,s'~N(0,1); a~U{0..A-1}; r~2N(0,1); done~Bern(p)."""
    s = rng.standard_normal((n, obs_dim), dtype=np.float32)
    a = rng.integers(0, num_actions, n, dtype=np.int64)
    r = (2.0 * rng.standard_normal(n)).astype(np.float32)
    s2 = rng.standard_normal((n, obs_dim), dtype=np.float32)
    d = rng.random(n) < done_p
    return s, a, r, s2, d
