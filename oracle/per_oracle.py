"""CPU restatement of the prioritized-replay sum tree (oracle; test infra only).  PARITY UNPINNED.

The reference has no prioritized replay (``General/Base/replay_buffer.py:68-85`` samples
uniformly), so there is nothing in the reference to pin this against: it restates
``deep-q-learning_b200/csrc/per.cu`` (proportional variant of Schaul et al. 2016) in numpy fp32 with the
same association (``parent = fl32(left + right)``, stratified ``u = (i + r) * (total / B)``, descend with
``u -= left`` when going right), so the kernels can be checked bit-for-bit.
"""
import numpy as np

from .philox import philox4x32_10

F32 = np.float32
PER_STREAM = 0x50455221


class OraclePER:
    def __init__(self, capacity, alpha=0.6, eps=1e-6, seed=0):
        self.capacity = int(capacity)
        self.L = 1
        while self.L < self.capacity:
            self.L *= 2
        self.levels = int(np.log2(self.L))
        self.alpha, self.eps, self.seed = F32(alpha), F32(eps), int(seed)
        self.tree = np.zeros(2 * self.L, dtype=F32)

    def rebuild(self):
        first = self.L // 2
        while first >= 1:
            self.tree[first:2 * first] = self.tree[2 * first:4 * first:2] + self.tree[2 * first + 1:4 * first:2]
            first //= 2

    def fill(self, prio):
        self.tree[self.L:] = 0
        self.tree[self.L:self.L + len(prio)] = np.asarray(prio, dtype=F32)
        self.rebuild()

    def update(self, idx, val, is_td=False):
        idx = np.asarray(idx, dtype=np.int64)
        val = np.asarray(val, dtype=F32)
        if is_td:
            val = np.power(np.abs(val) + self.eps, self.alpha).astype(F32)
        for i, v in zip(idx, val):                      # last write wins for duplicates, like ordered stores
            self.tree[self.L + i] = v
        nodes = self.L + idx
        for _ in range(self.levels):
            nodes = np.unique(nodes >> 1)
            self.tree[nodes] = self.tree[2 * nodes] + self.tree[2 * nodes + 1]

    def total(self):
        return self.tree[1]

    def sample(self, step, batch):
        i = np.arange(batch, dtype=np.uint64)
        o = philox4x32_10(i, int(step) & 0xFFFFFFFF, (int(step) >> 32) & 0xFFFFFFFF, PER_STREAM,
                          self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF)
        r = (o[0] >> np.uint32(8)).astype(F32) * F32(1.0 / 16777216.0)
        seg = F32(self.tree[1] / F32(batch))
        u = ((np.arange(batch, dtype=F32) + r) * seg).astype(F32)
        node = np.ones(batch, dtype=np.int64)
        for _ in range(self.levels):
            left = self.tree[2 * node]
            go_left = u < left
            u = np.where(go_left, u, (u - left).astype(F32)).astype(F32)
            node = np.where(go_left, 2 * node, 2 * node + 1)
        return node - self.L, self.tree[node]
