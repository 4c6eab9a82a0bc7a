"""Prioritized replay sampler (``dqn_per_*``): sum-tree proportional sampling and priority updates on the
device.  Extension beyond the reference (which samples uniformly); the indices it returns plug into
``Agent._step(indices=...)`` / ``DqnEngine.train_steps(indices=...)``."""
import ctypes as C

import numpy as np

from . import _lib


class PrioritizedSampler:
    def __init__(self, capacity, alpha=0.6, eps=1e-6, seed=0, device=0):
        import torch
        self.torch, self.lib = torch, _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DqnError(-3, "no CUDA device: the B200 DQN path has no CPU fallback")
        self.capacity, self.device = int(capacity), int(device)
        nbytes = C.c_uint64(0)
        _lib.check(self.lib.dqn_per_arena_bytes(self.capacity, C.byref(nbytes)))
        with torch.cuda.device(self.device):
            self._arena = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=f"cuda:{self.device}")
            base = (self._arena.data_ptr() + 255) & ~255
            stream = torch.cuda.current_stream(self.device).cuda_stream
        h = C.c_void_p()
        _lib.check(self.lib.dqn_per_create(self.device, self.capacity, float(alpha), float(eps), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                           stream, base, int(nbytes.value), C.byref(h)))
        self.h = h
        lb = C.c_int64(0)
        _lib.check(self.lib.dqn_per_leaf_base(self.h, C.byref(lb)))
        self.leaf_base = int(lb.value)
        levels = self.leaf_base.bit_length() - 1
        # kernels per priority update (csrc/per.cu): leaves, warp-per-update bottom 8 levels, full recompute of the top
        self.launches_per_update = 1 + levels if levels < 8 or levels > 28 else (3 if levels > 8 else 2)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dqn_per_destroy(self.h)
            self.h, self._arena = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, indices, values, is_td=False):
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        val = np.ascontiguousarray(values, dtype=np.float32)
        _lib.check(self.lib.dqn_per_update_host(self.h, _lib.ptr(idx), _lib.ptr(val), idx.size, 1 if is_td else 0))

    def update_device(self, idx_t, val_t, is_td=False):
        _lib.check(self.lib.dqn_per_update(self.h, C.c_void_p(idx_t.data_ptr()), C.c_void_p(val_t.data_ptr()), idx_t.numel(), 1 if is_td else 0))

    def fill_device(self, prio_t):
        _lib.check(self.lib.dqn_per_fill(self.h, C.c_void_p(prio_t.data_ptr()), prio_t.numel()))

    def sample(self, step, batch):
        idx, pr = np.empty(batch, np.int64), np.empty(batch, np.float32)
        _lib.check(self.lib.dqn_per_sample_host(self.h, int(step), int(batch), _lib.ptr(idx), _lib.ptr(pr)))
        return idx, pr

    def sample_device(self, step, idx_t, prio_t):
        _lib.check(self.lib.dqn_per_sample(self.h, int(step), idx_t.numel(), C.c_void_p(idx_t.data_ptr()), C.c_void_p(prio_t.data_ptr())))

    def total(self):
        out = C.c_float(0)
        _lib.check(self.lib.dqn_per_total(self.h, C.byref(out)))
        return float(out.value)

    def nodes(self, first, n):
        out = np.empty(n, np.float32)
        _lib.check(self.lib.dqn_per_read_nodes(self.h, int(first), int(n), _lib.ptr(out)))
        return out
