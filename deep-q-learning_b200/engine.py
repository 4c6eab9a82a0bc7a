"""``DqnEngine`` -- one ``dqn_handle`` (``include/dqn_b200.h``) with numpy-facing methods.

PyTorch is used for exactly two things: allocating the device arena (``torch.empty`` on the CUDA
device, whose ``data_ptr()`` is handed to ``dqn_create``) and naming the stream the kernels are
enqueued on (so ``torch.cuda.Event`` timing sees them).  All arithmetic is in ``libdqn_b200.so``.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .specs import flatten_tree, unflatten_tree, param_count


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class DqnEngine:
    def __init__(self, obs_dim, num_actions, buffer_size, batch_size, gamma, optimizer, n_agents=1,
                 seed=0, device=0, hidden=(32, 64), agent_id_base=0, step_kernel=None, session=False):
        import torch   # allocator + stream only

        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DqnError(-3, "no CUDA device: the B200 DQN path has no CPU fallback")
        self.torch = torch
        self.obs_dim, self.num_actions, self.n_agents = int(obs_dim), int(num_actions), int(n_agents)
        self.buffer_size, self.device = int(buffer_size), int(device)
        self.P = param_count(obs_dim, num_actions, hidden)
        cfg = _lib.DqnConfig()
        cfg.struct_size = C.sizeof(_lib.DqnConfig)
        cfg.device, cfg.n_agents = self.device, self.n_agents
        cfg.obs_dim, cfg.num_actions = self.obs_dim, self.num_actions
        cfg.hidden1, cfg.hidden2 = int(hidden[0]), int(hidden[1])
        cfg.batch_size, cfg.buffer_size, cfg.gamma = int(batch_size), self.buffer_size, float(gamma)
        cfg.opt_kind = _lib.DQN_OPT_ADAMW if optimizer.kind == "adamw" else _lib.DQN_OPT_ADAM
        cfg.lr, cfg.b1, cfg.b2 = optimizer.learning_rate, optimizer.b1, optimizer.b2
        cfg.eps, cfg.eps_root, cfg.weight_decay = optimizer.eps, optimizer.eps_root, optimizer.weight_decay
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.agent_id_base = int(agent_id_base)
        # which train-step kernel: "auto" (cluster while 4 * agents <= SMs), "cta" (one CTA per agent) or
        # "cluster" (one agent over a 4-CTA cluster); DQN_B200_STEP_KERNEL overrides the default
        step_kernel = step_kernel or os.environ.get("DQN_B200_STEP_KERNEL", "auto")
        cfg.step_kernel = _lib.STEP_KERNELS[step_kernel]
        self.step_kernel = step_kernel
        nbytes = C.c_uint64(0)
        _lib.check(self.lib.dqn_arena_bytes(C.byref(cfg), C.byref(nbytes)))
        with torch.cuda.device(self.device):
            self._arena = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=f"cuda:{self.device}")
            base = self._arena.data_ptr()
            cfg.arena = (base + 255) & ~255
            cfg.arena_bytes = int(nbytes.value)
            cfg.stream = torch.cuda.current_stream(self.device).cuda_stream
        self.arena_bytes = int(nbytes.value)
        h = C.c_void_p()
        _lib.check(self.lib.dqn_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._cfg = cfg
        if session:       # resident kernel serving add+step / act / sync commands from mapped host memory (single agent)
            _lib.check(self.lib.dqn_set_session(self.h, 1))
        self._loss1 = np.empty(1, np.float32)
        self._loss1_ptr = _lib.ptr(self._loss1)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.dqn_destroy(self.h)
            self.h = None
            self._arena = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_session(self, enable):
        _lib.check(self.lib.dqn_set_session(self.h, int(enable)))

    def synchronize(self):
        _lib.check(self.lib.dqn_synchronize(self.h))

    # -- parameters / optimiser state ----------------------------------------------------------------
    def set_params_flat(self, flat, agent=0, which=_lib.DQN_PARAMS_ONLINE):
        flat = _f32(flat)
        _lib.check(self.lib.dqn_set_params(self.h, agent, which, _lib.ptr(flat), flat.size))

    def get_params_flat(self, agent=0, which=_lib.DQN_PARAMS_ONLINE):
        out = np.empty(self.P, np.float32)
        _lib.check(self.lib.dqn_get_params(self.h, agent, which, _lib.ptr(out), out.size))
        return out

    def set_params(self, tree, agent=0, which=_lib.DQN_PARAMS_ONLINE):
        self.set_params_flat(flatten_tree(tree, self.obs_dim, self.num_actions), agent, which)

    def get_params(self, agent=0, which=_lib.DQN_PARAMS_ONLINE):
        return unflatten_tree(self.get_params_flat(agent, which), self.obs_dim, self.num_actions)

    def set_opt_state(self, count, mu_tree, nu_tree, agent=0):
        mu = flatten_tree(mu_tree, self.obs_dim, self.num_actions)
        nu = flatten_tree(nu_tree, self.obs_dim, self.num_actions)
        _lib.check(self.lib.dqn_set_opt_state(self.h, agent, int(count), _lib.ptr(mu), _lib.ptr(nu), mu.size))

    def get_opt_state(self, agent=0):
        mu, nu = np.empty(self.P, np.float32), np.empty(self.P, np.float32)
        cnt = C.c_int32(0)
        _lib.check(self.lib.dqn_get_opt_state(self.h, agent, C.byref(cnt), _lib.ptr(mu), _lib.ptr(nu), self.P))
        return (np.int32(cnt.value), unflatten_tree(mu, self.obs_dim, self.num_actions),
                unflatten_tree(nu, self.obs_dim, self.num_actions))

    def set_hparams(self, agent=0, gamma=-1.0, batch_size=-1, lr=-1.0, b1=-1.0, b2=-1.0, eps=-1.0,
                    eps_root=-1.0, weight_decay=-1.0):
        hp = _lib.DqnHparams(gamma, batch_size, lr, b1, b2, eps, eps_root, weight_decay)
        _lib.check(self.lib.dqn_set_hparams(self.h, agent, C.byref(hp)))

    def get_hparams(self, agent=0):
        hp = _lib.DqnHparams()
        _lib.check(self.lib.dqn_get_hparams(self.h, agent, C.byref(hp)))
        return {k: getattr(hp, k) for k, _ in _lib.DqnHparams._fields_}

    # -- whole-agent state for resume (what the reference's checkpoint omits: target net, ring, counters) --------------
    def get_counters(self, agent=0):
        c = _lib.DqnCounters()
        _lib.check(self.lib.dqn_get_counters(self.h, agent, C.byref(c)))
        return {k: getattr(c, k) for k, _ in _lib.DqnCounters._fields_ if k != "reserved"}

    def set_counters(self, counters, agent=0):
        c = _lib.DqnCounters(int(counters["ring_counter"]), int(counters["train_steps"]), int(counters["adam_count"]), 0,
                             float(counters["pb1"]), float(counters["pb2"]))
        _lib.check(self.lib.dqn_set_counters(self.h, agent, C.byref(c)))

    def export_state(self, agent=0):
        """Everything needed to continue this agent bit for bit (numpy arrays / scalars, ``np.savez``-able)."""
        cnt, mu, nu = np.int32(0), np.empty(self.P, np.float32), np.empty(self.P, np.float32)
        c = C.c_int32(0)
        _lib.check(self.lib.dqn_get_opt_state(self.h, agent, C.byref(c), _lib.ptr(mu), _lib.ptr(nu), self.P))
        s, a, r, s2, d = self.buffer_export(agent)
        out = dict(params=self.get_params_flat(agent, _lib.DQN_PARAMS_ONLINE), target_params=self.get_params_flat(agent, _lib.DQN_PARAMS_TARGET),
                   mu=mu, nu=nu, states=s, actions=a, rewards=r, observations=s2, dones=d)
        out.update({"counter_" + k: np.asarray(v) for k, v in self.get_counters(agent).items()})
        out.update({"hparam_" + k: np.asarray(v) for k, v in self.get_hparams(agent).items()})
        return out

    def import_state(self, state, agent=0):
        hp = {k[7:]: state[k].item() for k in state if k.startswith("hparam_")}
        self.set_hparams(agent, **hp)
        self.set_params_flat(state["params"], agent, _lib.DQN_PARAMS_ONLINE)
        self.set_params_flat(state["target_params"], agent, _lib.DQN_PARAMS_TARGET)
        counters = {k[8:]: state[k].item() for k in state if k.startswith("counter_")}
        mu, nu = _f32(state["mu"]), _f32(state["nu"])
        _lib.check(self.lib.dqn_set_opt_state(self.h, agent, int(counters["adam_count"]), _lib.ptr(mu), _lib.ptr(nu), mu.size))
        self.set_counters(dict(counters, ring_counter=0), agent)                    # slots are restored in slot order from 0 ...
        self.store(state["states"], state["actions"], state["rewards"], state["observations"], state["dones"], agent=agent)
        self.set_counters(counters, agent)                                           # ... then the real counters (and decay powers)

    # -- replay ring ----------------------------------------------------------------------------------
    def store(self, s, a, r, s2, done, agent=0):
        s, s2, r = _f32(s), _f32(s2), _f32(r)
        a = np.ascontiguousarray(a, dtype=np.int64)
        done = np.ascontiguousarray(done, dtype=np.bool_)
        n = a.shape[0]
        if s.shape != (n, self.obs_dim) or s2.shape != (n, self.obs_dim) or r.shape != (n,) or done.shape != (n,):
            raise ValueError("store: inconsistent transition array shapes")
        _lib.check(self.lib.dqn_store(self.h, agent, n, _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(s2), _lib.ptr(done)))

    def buffer_state(self, agent=0):
        size, counter = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.dqn_buffer_state(self.h, agent, C.byref(size), C.byref(counter)))
        return int(size.value), int(counter.value)

    def _batch_arrays(self, n):
        return (np.empty((n, self.obs_dim), np.float32), np.empty(n, np.int64), np.empty(n, np.float32),
                np.empty((n, self.obs_dim), np.float32), np.empty(n, np.bool_))

    def buffer_export(self, agent=0):
        s, a, r, s2, d = self._batch_arrays(self.buffer_size)
        _lib.check(self.lib.dqn_buffer_export(self.h, agent, _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(s2), _lib.ptr(d)))
        return s, a, r, s2, d

    def sample_indices(self, step, batch_size, agent=0):
        out = np.empty(batch_size, np.int64)
        _lib.check(self.lib.dqn_sample_indices(self.h, agent, int(step), int(batch_size), _lib.ptr(out)))
        return out

    def sample_batch(self, batch_size, indices=None, step=0, agent=0):
        s, a, r, s2, d = self._batch_arrays(batch_size)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
        if idx is not None and idx.shape != (batch_size,):
            raise ValueError("sample_batch: indices must have shape (batch_size,)")
        _lib.check(self.lib.dqn_sample_batch(self.h, agent, _lib.ptr(idx), int(step), int(batch_size),
                                             _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(s2), _lib.ptr(d)))
        return s, a, r, s2, d

    # -- train step / target sync / acting ---------------------------------------------------------
    def train_steps(self, K=1, indices=None, agent_begin=0, agent_end=None):
        agent_end = self.n_agents if agent_end is None else agent_end
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
        _lib.check(self.lib.dqn_train_step(self.h, agent_begin, agent_end, int(K), _lib.ptr(idx), None))

    def train_step_debug(self, indices=None, agent=0):
        """One train step with every intermediate copied back (parity tests)."""
        hp = self.get_hparams(agent)
        B, A = hp["batch_size"], self.num_actions
        out = dict(indices=np.empty(B, np.int64), q=np.empty((B, A), np.float32),
                   next_q=np.empty((B, A), np.float32), next_q_tm=np.empty((B, A), np.float32),
                   max_actions=np.empty(B, np.int32), targets=np.empty((B, A), np.float32),
                   loss=np.empty(1, np.float32), grads=np.empty(self.P, np.float32))
        taps = _lib.DqnDebugTaps(*[_lib.ptr(out[k]) for k, _ in _lib.DqnDebugTaps._fields_])
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
        _lib.check(self.lib.dqn_train_step(self.h, agent, agent + 1, 1, _lib.ptr(idx), C.byref(taps)))
        out["loss"] = out["loss"][0]
        out["grads_flat"] = out["grads"]
        out["grads"] = unflatten_tree(out["grads"], self.obs_dim, self.num_actions)
        return out

    def losses(self, n, agent=0):
        out = np.empty(n, np.float32)
        ts = C.c_int64(0)
        _lib.check(self.lib.dqn_get_losses(self.h, agent, int(n), _lib.ptr(out), C.byref(ts)))
        return out

    def last_loss(self, agent=0):
        """Loss of the most recent train step (waits for it; the kernel writes it to mapped host memory)."""
        _lib.check(self.lib.dqn_get_losses(self.h, agent, 1, self._loss1_ptr, None))
        return float(self._loss1[0])

    def lagged_loss(self, agent=0, lag=1):
        """Loss of the train step ``lag`` steps before the most recent one; waits for that step only (in session mode the
        most recent step may still be in flight)."""
        _lib.check(self.lib.dqn_get_loss_lagged(self.h, agent, int(lag), self._loss1_ptr))
        return float(self._loss1[0])

    def train_step_count(self, agent=0):
        ts = C.c_int64(0)
        _lib.check(self.lib.dqn_get_losses(self.h, agent, 0, None, C.byref(ts)))
        return int(ts.value)

    def sync_target(self, agent_begin=0, agent_end=None):
        agent_end = self.n_agents if agent_end is None else agent_end
        _lib.check(self.lib.dqn_sync_target(self.h, agent_begin, agent_end))

    def polyak_target(self, tau, agent_begin=0, agent_end=None):
        """Soft target update theta^- := tau theta + (1 - tau) theta^- (extension; the reference hard-copies)."""
        agent_end = self.n_agents if agent_end is None else agent_end
        _lib.check(self.lib.dqn_polyak_target(self.h, agent_begin, agent_end, float(tau)))

    def set_loss(self, kind, agent_begin=0, agent_end=None):
        """``"huber"`` (the reference's loss, default) or ``"l2"`` / ``"mse"`` (0.5 e^2, extension)."""
        agent_end = self.n_agents if agent_end is None else agent_end
        _lib.check(self.lib.dqn_set_loss_kind(self.h, agent_begin, agent_end, _lib.LOSS_KINDS[kind]))

    # -- episode-loop control on the device (torch CUDA tensors in, torch CUDA tensors out; enqueue only) ----------
    def configure_episodes(self, configs, agent_begin=0, reset_counters=True):
        """``configs``: one dict per agent with the reference's kwarg names (epsilon, epsilon_decay_rate, min_epsilon,
        reward_to_reach, max_episodes, max_steps, training_start, train_frequency, replace_frequency)."""
        arr = (_lib.DqnEpisodeConfig * len(configs))()
        for c, src in zip(arr, configs):
            for k, _ in _lib.DqnEpisodeConfig._fields_:
                if k != "reserved":
                    setattr(c, k, src[k])
        _lib.check(self.lib.dqn_episode_configure(self.h, agent_begin, agent_begin + len(configs), arr, 1 if reset_counters else 0))

    def _dev(self, t, dtype, n):
        torch = self.torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous() and t.numel() == n):
            raise TypeError(f"expected a contiguous CUDA tensor of {n} x {dtype}")
        return C.c_void_p(t.data_ptr())

    def policy(self, states, actions_out=None, agent_begin=0, agent_end=None):
        """Agent._policy for agents [agent_begin, agent_end): states f32[n, D] -> actions i32[n] (epsilon-greedy)."""
        torch = self.torch
        agent_end = self.n_agents if agent_end is None else agent_end
        n = agent_end - agent_begin
        if actions_out is None:
            actions_out = torch.empty(n, dtype=torch.int32, device=states.device)
        _lib.check(self.lib.dqn_policy_batch(self.h, agent_begin, agent_end, self._dev(states, torch.float32, n * self.obs_dim),
                                             self._dev(actions_out, torch.int32, n)))
        return actions_out

    def observe(self, states, actions, rewards, observations, dones, episode_end_out=None, agent_begin=0, agent_end=None):
        """One iteration of Agent._run_episode after env.step for every agent of the range; returns u8[n] episode-end flags."""
        torch = self.torch
        agent_end = self.n_agents if agent_end is None else agent_end
        n = agent_end - agent_begin
        if episode_end_out is None:
            episode_end_out = torch.empty(n, dtype=torch.uint8, device=states.device)
        _lib.check(self.lib.dqn_observe_batch(
            self.h, agent_begin, agent_end, self._dev(states, torch.float32, n * self.obs_dim), self._dev(actions, torch.int32, n),
            self._dev(rewards, torch.float32, n), self._dev(observations, torch.float32, n * self.obs_dim),
            self._dev(dones, torch.uint8, n), self._dev(episode_end_out, torch.uint8, n)))
        return episode_end_out

    def train_flagged(self, agent_begin=0, agent_end=None):
        agent_end = self.n_agents if agent_end is None else agent_end
        _lib.check(self.lib.dqn_train_flagged(self.h, agent_begin, agent_end))

    def episode_state(self, agent=0):
        st = _lib.DqnEpisodeState()
        _lib.check(self.lib.dqn_episode_get_state(self.h, agent, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _lib.DqnEpisodeState._fields_}

    def act(self, state, agent=0):
        st = _f32(state).reshape(-1)
        if st.size != self.obs_dim:
            raise ValueError(f"act: state has {st.size} entries, expected {self.obs_dim}")
        out = C.c_int32(0)
        _lib.check(self.lib.dqn_act(self.h, agent, _lib.ptr(st), C.byref(out)))
        return int(out.value)

    def act_batch(self, states, agent_begin=0, agent_end=None):
        agent_end = self.n_agents if agent_end is None else agent_end
        st = _f32(states).reshape(agent_end - agent_begin, self.obs_dim)
        out = np.empty(agent_end - agent_begin, np.int32)
        _lib.check(self.lib.dqn_act_batch(self.h, agent_begin, agent_end, _lib.ptr(st), _lib.ptr(out)))
        return out
