"""Prioritized-replay kernels vs the numpy oracle (bit-exact: tree nodes, sampled indices and priorities), and
size-independent properties at the BASELINE configs[4] size (16M leaves, batch 32768)."""
import numpy as np
import pytest

import dqn_b200
from oracle.per_oracle import OraclePER

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cap", [2, 100, 1000, 4096])
def test_tree_and_samples_bit_exact_vs_oracle(cap):
    rng = np.random.default_rng(cap)
    per = dqn_b200.PrioritizedSampler(cap, alpha=0.6, eps=1e-6, seed=11)
    ora = OraclePER(cap, alpha=0.6, eps=1e-6, seed=11)
    for rnd in range(5):
        n = int(rng.integers(1, min(cap, 300) + 1))
        idx = rng.integers(0, cap, n)
        if rnd % 2 == 0:
            idx = np.unique(idx)                                   # duplicate-free: raw priorities
            val = rng.random(idx.size).astype(np.float32) + 0.01
            per.update(idx, val)
            ora.update(idx, val)
        else:                                                       # TD errors -> (|td| + eps)^alpha; duplicates carry equal values
            td = (rng.standard_normal(cap) * 2).astype(np.float32)
            per.update(idx, td[idx], is_td=True)
            ora.update(idx, td[idx], is_td=True)
        got = per.nodes(0, 2 * ora.L)
        # powf on the device may differ from numpy's by an ulp: compare leaves within 2 ulp, then sums exactly from device leaves
        leaves = got[ora.L:]
        np.testing.assert_allclose(leaves, ora.tree[ora.L:], rtol=3e-7, atol=0)
        ora.tree[ora.L:] = leaves
        ora.rebuild()
        assert np.array_equal(got[1:], ora.tree[1:]), f"cap {cap} round {rnd}"
        for step in (0, 7):
            B = int(rng.integers(1, 200))
            gi, gp = per.sample(step, B)
            oi, op = ora.sample(step, B)
            assert np.array_equal(gi, oi) and np.array_equal(gp, op)
            assert gi.min() >= 0 and gi.max() < cap and np.all(gp > 0)
    assert per.total() == float(ora.total())


def test_full_size_properties():
    import torch
    cap, B = 16 * 2**20, 32768
    per = dqn_b200.PrioritizedSampler(cap, seed=5)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    prio = torch.rand(cap, generator=g, device="cuda") + 1e-3
    prio[: cap // 2] *= 3.0                                         # first half three times as likely
    per.fill_device(prio)
    total = per.total()
    assert abs(total - float(prio.double().sum())) / total < 1e-4   # fp32 tree sum vs fp64 sum
    idx = torch.empty(B, dtype=torch.int64, device="cuda"); pr = torch.empty(B, dtype=torch.float32, device="cuda")
    first_half = 0
    for step in range(20):
        per.sample_device(step, idx, pr)
        assert int(idx.min()) >= 0 and int(idx.max()) < cap
        assert torch.equal(pr, prio[idx])                           # returned priority is the leaf's priority
        first_half += int((idx < cap // 2).sum())
    assert abs(first_half / (20 * B) - 0.75) < 0.01                 # P(first half) = 3/(3+1)
    # update -> sampled leaves get priority 0 except one: all mass moves there
    per.sample_device(99, idx, pr)
    per.update_device(idx, torch.zeros(B, device="cuda"))
    lo = per.total()
    assert lo < total
    # idempotence: re-applying the same update leaves the tree unchanged
    root_path = per.nodes(1, 1)[0]
    per.update_device(idx, torch.zeros(B, device="cuda"))
    assert per.nodes(1, 1)[0] == root_path
