"""Drop-in for ``General/Base/replay_buffer.py`` over the HBM-resident ring.

Same names and argument meaning as the reference: ``ReplayBuffer(buffer_size, obs_shape, ac_shape)``
(``replay_buffer.py:20``), ``.add(state, action, reward, observation, done)`` (``:58``), ``.size``
and the five array properties (``:34-56``), and the free function ``sample_batch(num_samples,
states, actions, rewards, observations, dones, batch_size)`` (``:68-85``).  The arrays live on the
device as AoS records; the properties return ``RingView`` handles that ``sample_batch`` recognises
and that convert to the reference's numpy arrays (dtype and contents bit-identical) on
``numpy.asarray``.
"""
import numpy as np

from . import _lib
from .specs import adam

try:                       # optional CPython accelerator of the host-side staging in add() (csrc/hoststage.c); no device code
    from . import _hoststage
except ImportError:        # not built: the numpy stores below do the same thing
    _hoststage = None

_FIELDS = ("states", "actions", "rewards", "observations", "dones")


class RingView:
    """Lazy view of one of the five replay arrays (device-resident)."""

    def __init__(self, buffer, field):
        self._buffer, self._field = buffer, field

    def __array__(self, dtype=None, copy=None):
        arr = self._buffer._export()[_FIELDS.index(self._field)]
        return arr if dtype is None else arr.astype(dtype)

    def __getitem__(self, item):
        return np.asarray(self)[item]

    def __len__(self):
        return self._buffer._buffer_size

    @property
    def shape(self):
        n, d = self._buffer._buffer_size, self._buffer._obs_dim
        return (n, d) if self._field in ("states", "observations") else (n,)

    @property
    def dtype(self):
        return {"states": np.dtype(np.float32), "observations": np.dtype(np.float32),
                "rewards": np.dtype(np.float32), "actions": np.dtype(np.int64),
                "dones": np.dtype(np.bool_)}[self._field]


class ReplayBuffer:
    def __init__(self, buffer_size, obs_shape, ac_shape, _engine=None, _agent=0, device=0, seed=0):
        obs_shape, ac_shape = tuple(obs_shape), tuple(ac_shape)
        if len(obs_shape) != 2 or obs_shape[0] != buffer_size or ac_shape != (buffer_size,):
            raise ValueError("obs_shape must be (buffer_size, D) and ac_shape (buffer_size,) "
                             "(Test/lunar_lander.py:40-41)")
        self._buffer_size = int(buffer_size)
        self._obs_dim = int(obs_shape[1])
        if _engine is None:
            from .engine import DqnEngine
            _engine = DqnEngine(self._obs_dim, 4, self._buffer_size, 1, 0.0, adam(0.0), device=device, seed=seed)
        self._engine, self._agent = _engine, _agent
        d = self._obs_dim
        # host staging for add(): transitions are appended here and moved to the device ring by ONE
        # coalesced store right before anything can observe the ring (train step, sample, export, size
        # of the device counter).  Semantically identical to n scalar adds; it turns the reference's
        # 4 adds per train step (train_frequency, Test/lunar_lander.py:30) into one H2D copy + one launch.
        self._cap = 64
        self._ps, self._po = np.zeros((self._cap, d), np.float32), np.zeros((self._cap, d), np.float32)
        self._pa, self._pr, self._pd = np.zeros(self._cap, np.int64), np.zeros(self._cap, np.float32), np.zeros(self._cap, np.bool_)
        self._ptrs = tuple(_lib.ptr(x) for x in (self._ps, self._pa, self._pr, self._po, self._pd))
        self._stage = _hoststage.new(*self._ptrs, d, self._cap) if _hoststage is not None else None
        self._bound = False
        if self._stage is not None and hasattr(_hoststage, "bind"):
            # the per-step C-ABI calls by address (no ctypes marshalling): dqn_store_train_step / dqn_get_losses of THIS handle
            import ctypes as C
            lib = _engine.lib
            _hoststage.bind(self._stage, int(_engine.h.value), C.cast(lib.dqn_store_train_step, C.c_void_p).value,
                            C.cast(lib.dqn_get_losses, C.c_void_p).value, C.cast(lib.dqn_get_loss_lagged, C.c_void_p).value, int(_agent))
            self._bound = True
        self._pending = 0
        self._counter = 0
        self._num_samples = 0
        self._sample_calls = 0

    # -- reference surface ----------------------------------------------------------------------------
    @property
    def size(self):
        return self._num_samples

    states = property(lambda self: RingView(self, "states"))
    actions = property(lambda self: RingView(self, "actions"))
    rewards = property(lambda self: RingView(self, "rewards"))
    observations = property(lambda self: RingView(self, "observations"))
    dones = property(lambda self: RingView(self, "dones"))

    def add(self, state, action, reward, observation, done):
        i = self._pending
        if self._stage is not None:
            try:
                _hoststage.put(self._stage, i, state, action, reward, observation, done)   # the five stores, in C
            except (TypeError, BufferError, ValueError):   # not C-contiguous float32 rows / plain scalars: let numpy convert
                self._add_numpy(i, state, action, reward, observation, done)
        else:
            self._add_numpy(i, state, action, reward, observation, done)
        self._pending = i + 1
        self._counter += 1
        self._num_samples = min(self._counter, self._buffer_size)
        if self._pending == self._cap:
            self.flush()

    def _add_numpy(self, i, state, action, reward, observation, done):
        self._ps[i] = state
        self._pa[i] = action
        self._pr[i] = reward
        self._po[i] = observation
        self._pd[i] = done

    def step_staged(self):
        """``dqn_store_train_step(h, agent, n_staged, <staging arrays>, K=1, NULL)``: the staged add()s and ONE train step as one
        command / launch.  Returns after enqueueing."""
        n = self._pending
        e = self._engine
        if self._bound:
            rc = _hoststage.step(self._stage, n)
            if rc:
                _lib.check(rc)
        else:
            _lib.check(e.lib.dqn_store_train_step(e.h, self._agent, n, *self._ptrs, 1, None))
        self._pending = 0                                   # only once the library has taken the staged transitions

    def last_loss(self, lag=0):
        """Loss of the most recent train step of this buffer's agent (waits for it).  ``lag=1``: the loss of the step BEFORE the
        most recent one, waiting for that step only -- with ``Agent(session=True)`` the most recent step may still be in flight,
        so a loop that calls ``_step()`` and then ``last_loss(1)`` reads every loss, one step behind, without stalling the device."""
        if self._bound:
            rc, loss = _hoststage.loss(self._stage, lag)
            if rc:
                _lib.check(rc)
            return loss
        return self._engine.lagged_loss(self._agent, lag) if lag else self._engine.last_loss(self._agent)

    def flush(self):
        """Move staged add() transitions into the device ring (no-op when nothing is pending)."""
        n = self._pending
        if n:
            e = self._engine
            _lib.check(e.lib.dqn_store(e.h, self._agent, n, *self._ptrs))
            self._pending = 0

    # -- vectorised extension ------------------------------------------------------------------------
    def add_many(self, states, actions, rewards, observations, dones):
        """n ``add`` calls in order, as one coalesced device store."""
        self.flush()
        self._engine.store(states, actions, rewards, observations, dones, agent=self._agent)
        self._counter += len(np.asarray(actions))
        self._num_samples = min(self._counter, self._buffer_size)

    def sample(self, batch_size, indices=None):
        self.flush()
        step = self._sample_calls
        self._sample_calls += 1
        return self._engine.sample_batch(batch_size, indices=indices, step=step, agent=self._agent)

    def _export(self):
        self.flush()
        return self._engine.buffer_export(self._agent)


def sample_batch(num_samples, states, actions, rewards, observations, dones, batch_size, indices=None):
    """``sample_batch`` of the reference, on device-resident arrays.

    The five array arguments must be the ``RingView`` properties of one ``ReplayBuffer``;
    ``num_samples`` must be its ``size`` (the reference passes ``replay_buffer.size``,
    ``q_agent.py:147``).  Host numpy arrays are rejected: there is no CPU path.
    """
    views = (states, actions, rewards, observations, dones)
    if not all(isinstance(v, RingView) for v in views):
        raise TypeError("sample_batch: pass the ReplayBuffer's own array properties (device-resident); "
                        "the B200 path has no CPU gather")
    buf = states._buffer
    if any(v._buffer is not buf for v in views) or tuple(v._field for v in views) != _FIELDS:
        raise ValueError("sample_batch: arrays must be (states, actions, rewards, observations, dones) of one buffer")
    if num_samples != buf.size:
        raise ValueError("sample_batch: num_samples must equal replay_buffer.size")
    return buf.sample(batch_size, indices=indices)
