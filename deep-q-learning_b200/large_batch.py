"""Large-batch data-parallel DDQN step (BASELINE configs[3]) over ``dqn_lb_*`` (``include/dqn_b200.h``).

One agent, global minibatch ``B`` split contiguously over ``world_size`` ranks (one process per GPU).
Per step every rank runs forward + backward on its ``B / world_size`` rows with gradients pre-scaled by
``1/B`` (the loss of ``q_learning_functions.py:36`` is a mean over the batch, so shard gradients add up),
the flat gradient (+ the loss, riding in the last slot) is summed, and every rank applies the identical Adam
update -- replicas stay bit-identical because they all consume the same reduced buffer.

The sum is the library's own kernel over NVLink peer memory (``collective="p2p"``, ``csrc/comm_p2p.cu``: flag
barrier -> each rank reduces its slice of every rank's gradient window in rank order -> stores it into every
window -> flag barrier; deterministic) for 2, 4 or 8 ranks of one node, with ``torch.distributed.all_reduce``
(NCCL) as the comparison path (``collective="nccl"``).  torch is used for the arena allocation, for exchanging
the 64-byte IPC handles at start-up and for the NCCL comparison; all arithmetic is in ``libdqn_b200.so``.
"""
import ctypes as C

import numpy as np

from . import _lib
from .specs import flatten_tree, unflatten_tree, param_count

GEMM_MODES = {"fp32": 0, "tc3xtf32": 1}


class _DeviceFloats:
    """A raw device pointer as a CUDA-array-interface object (so torch can view library-owned memory)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class LargeBatchTrainer:
    def __init__(self, obs_dim, num_actions, hidden, batch_global, buffer_size, gamma, optimizer, rank=0,
                 world_size=1, seed=0, device=0, gemm_mode="fp32", process_group=None, collective="auto",
                 connect=True):
        import torch
        self.torch, self.lib = torch, _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DqnError(-3, "no CUDA device: the B200 DQN path has no CPU fallback")
        if batch_global % world_size:
            raise ValueError("batch_global must be divisible by world_size")
        self.obs_dim, self.num_actions, self.hidden = int(obs_dim), int(num_actions), tuple(hidden)
        self.rank, self.world, self.device, self.pg = int(rank), int(world_size), int(device), process_group
        self.batch_global, self.batch_local = int(batch_global), int(batch_global) // int(world_size)
        self.P = param_count(obs_dim, num_actions, self.hidden)
        cfg = _lib.DqnLbConfig()
        cfg.struct_size = C.sizeof(_lib.DqnLbConfig)
        cfg.device, cfg.obs_dim, cfg.num_actions = self.device, self.obs_dim, self.num_actions
        cfg.hidden1, cfg.hidden2, cfg.batch_local = int(hidden[0]), int(hidden[1]), self.batch_local
        cfg.gemm_mode, cfg.buffer_size, cfg.gamma = GEMM_MODES[gemm_mode], int(buffer_size), float(gamma)
        cfg.opt_kind = _lib.DQN_OPT_ADAMW if optimizer.kind == "adamw" else _lib.DQN_OPT_ADAM
        cfg.lr, cfg.b1, cfg.b2 = optimizer.learning_rate, optimizer.b1, optimizer.b2
        cfg.eps, cfg.eps_root, cfg.weight_decay = optimizer.eps, optimizer.eps_root, optimizer.weight_decay
        cfg.seed, cfg.rank, cfg.world = int(seed) & 0xFFFFFFFFFFFFFFFF, self.rank, self.world
        nbytes = C.c_uint64(0)
        _lib.check(self.lib.dqn_lb_arena_bytes(C.byref(cfg), C.byref(nbytes)))
        with torch.cuda.device(self.device):
            self._arena = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=f"cuda:{self.device}")
            base = (self._arena.data_ptr() + 255) & ~255
            cfg.arena, cfg.arena_bytes = base, int(nbytes.value)
            cfg.stream = torch.cuda.current_stream(self.device).cuda_stream
        self.arena_bytes = int(nbytes.value)
        h = C.c_void_p()
        _lib.check(self.lib.dqn_lb_create(C.byref(cfg), C.byref(h)))
        self.h = h
        ptr, cnt = C.c_void_p(), C.c_int64(0)
        _lib.check(self.lib.dqn_lb_grads(self.h, C.byref(ptr), C.byref(cnt)))
        off = ptr.value - self._arena.data_ptr()
        self.grads = self._arena[off:off + 4 * cnt.value].view(torch.float32)     # P gradients + loss, for the collective
        if collective == "auto":
            collective = "p2p" if self.world in (2, 4, 8) else "nccl"
        self.collective = collective if self.world > 1 else "none"
        self.window = None
        if self.collective == "p2p":
            handle = np.zeros(64, np.uint8)
            win = C.c_void_p()
            _lib.check(self.lib.dqn_lb_comm_init(self.h, _lib.ptr(handle), C.byref(win)))
            self.window = win.value
            with torch.cuda.device(self.device):                                   # the gradient now lives in the window
                self.grads = torch.as_tensor(_DeviceFloats(win.value, cnt.value), device=f"cuda:{self.device}")
            if connect:                                                            # one process per GPU: swap IPC handles
                import torch.distributed as dist
                mine = torch.from_numpy(handle).to(f"cuda:{self.device}")
                every = torch.empty(64 * self.world, dtype=torch.uint8, device=f"cuda:{self.device}")
                dist.all_gather_into_tensor(every, mine, group=self.pg)
                handles = np.ascontiguousarray(every.cpu().numpy())
                _lib.check(self.lib.dqn_lb_comm_connect(self.h, _lib.ptr(handles), None))

    @staticmethod
    def connect_in_process(trainers):
        """Ranks that live in ONE process (tests; several handles per GPU or one per visible GPU): exchange raw
        window pointers instead of IPC handles.  Each trainer must enqueue on its own stream."""
        world = len(trainers)
        wins = (C.c_void_p * world)(*[t.window for t in sorted(trainers, key=lambda t: t.rank)])
        for t in trainers:
            _lib.check(t.lib.dqn_lb_comm_connect(t.h, None, wins))

    def close(self):
        if getattr(self, "h", None):
            self.lib.dqn_lb_destroy(self.h)
            self.h, self._arena, self.grads = None, None, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ----------------------------------------------------------------------------------------
    def set_params(self, tree, which=0):
        flat = flatten_tree(tree, self.obs_dim, self.num_actions, self.hidden)
        _lib.check(self.lib.dqn_lb_set_params(self.h, which, _lib.ptr(flat), flat.size))

    def get_params(self, which=0):
        out = np.empty(self.P, np.float32)
        _lib.check(self.lib.dqn_lb_get_params(self.h, which, _lib.ptr(out), out.size))
        return unflatten_tree(out, self.obs_dim, self.num_actions, self.hidden)

    def set_opt_state(self, count, mu_tree, nu_tree):
        mu = flatten_tree(mu_tree, self.obs_dim, self.num_actions, self.hidden)
        nu = flatten_tree(nu_tree, self.obs_dim, self.num_actions, self.hidden)
        _lib.check(self.lib.dqn_lb_set_opt_state(self.h, int(count), _lib.ptr(mu), _lib.ptr(nu), mu.size))

    def get_opt_state(self):
        mu, nu, cnt = np.empty(self.P, np.float32), np.empty(self.P, np.float32), C.c_int32(0)
        _lib.check(self.lib.dqn_lb_get_opt_state(self.h, C.byref(cnt), _lib.ptr(mu), _lib.ptr(nu), self.P))
        return (np.int32(cnt.value), unflatten_tree(mu, self.obs_dim, self.num_actions, self.hidden),
                unflatten_tree(nu, self.obs_dim, self.num_actions, self.hidden))

    def store(self, s, a, r, s2, done):
        s, s2, r = (np.ascontiguousarray(x, dtype=np.float32) for x in (s, s2, r))
        a = np.ascontiguousarray(a, dtype=np.int64)
        done = np.ascontiguousarray(done, dtype=np.bool_)
        _lib.check(self.lib.dqn_lb_store(self.h, a.shape[0], _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(s2), _lib.ptr(done)))

    def store_device(self, s, a, r, s2, done):
        """Device tensors (torch): f32[n,D], i64[n], f32[n], f32[n,D], u8[n]."""
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.dqn_lb_store_device(self.h, a.shape[0], p(s), p(a), p(r), p(s2), p(done)))

    # -- one train step ---------------------------------------------------------------------------------
    def forward_backward(self, indices=None, debug=False):
        idx = self._local_indices(indices)
        _lib.check(self.lib.dqn_lb_forward_backward(self.h, _lib.ptr(idx), 1 if debug else 0))

    def all_reduce(self):
        if self.collective == "p2p":
            _lib.check(self.lib.dqn_lb_allreduce(self.h))      # enqueue only; every rank must call it once per step
        elif self.collective == "nccl":
            import torch.distributed as dist
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.pg)

    def apply(self):
        _lib.check(self.lib.dqn_lb_apply(self.h))

    def _local_indices(self, indices):
        if indices is None:
            return None
        idx = np.ascontiguousarray(indices, dtype=np.int64)      # global index list of the step; this rank takes its contiguous slice
        if idx.shape == (self.batch_global,):
            idx = np.ascontiguousarray(idx[self.rank * self.batch_local:(self.rank + 1) * self.batch_local])
        elif idx.shape != (self.batch_local,):
            raise ValueError("indices must have batch_global or batch_local entries")
        return idx

    def step(self, indices=None):
        """One train step.  With the library's own exchange (or a single rank) this is ONE C call that overlaps the exchange
        of the W2 gradient with the rest of backward; with NCCL it is forward_backward -> all_reduce -> apply."""
        if self.collective in ("p2p", "none"):
            _lib.check(self.lib.dqn_lb_train_step(self.h, _lib.ptr(self._local_indices(indices)), 0))
        else:
            self.forward_backward(indices)
            self.all_reduce()
            self.apply()

    def sync_target(self):
        _lib.check(self.lib.dqn_lb_sync_target(self.h))

    def polyak_target(self, tau):
        _lib.check(self.lib.dqn_lb_polyak_target(self.h, float(tau)))

    def set_loss(self, kind):
        _lib.check(self.lib.dqn_lb_set_loss_kind(self.h, _lib.LOSS_KINDS[kind]))

    def loss(self):
        out = C.c_float(0)
        _lib.check(self.lib.dqn_lb_get_loss(self.h, C.byref(out)))
        return float(out.value)

    def synchronize(self):
        _lib.check(self.lib.dqn_lb_synchronize(self.h))

    def debug_read_activations(self):
        """Hidden activations h1, h2 of the (theta, s) rows of the last forward_backward (numpy) -- parity tests use their
        signs to follow the device's branch of relu' where a pre-activation is within fp32 round-off of zero."""
        h1 = np.empty((self.batch_local, self.hidden[0]), np.float32)
        h2 = np.empty((self.batch_local, self.hidden[1]), np.float32)
        _lib.check(self.lib.dqn_lb_debug_read(self.h, 5, _lib.ptr(h1), h1.nbytes))
        _lib.check(self.lib.dqn_lb_debug_read(self.h, 6, _lib.ptr(h2), h2.nbytes))
        return h1, h2

    def debug_read(self):
        """Intermediates of the last forward_backward(debug=True) as numpy (parity tests)."""
        B, A = self.batch_local, self.num_actions
        q = np.empty((3 * B, A), np.float32)
        tg, ma = np.empty((B, A), np.float32), np.empty(B, np.int32)
        g, ix = np.empty(self.P + 1, np.float32), np.empty(B, np.int64)
        for what, arr in ((0, q), (1, tg), (2, ma), (3, g), (4, ix)):
            _lib.check(self.lib.dqn_lb_debug_read(self.h, what, _lib.ptr(arr), arr.nbytes))
        return dict(q=q[:B], next_q=q[B:2 * B], next_q_tm=q[2 * B:], targets=tg, max_actions=ma, indices=ix,
                    loss=float(g[-1]), grads_flat=g[:-1],
                    grads=unflatten_tree(g[:-1], self.obs_dim, self.num_actions, self.hidden))
