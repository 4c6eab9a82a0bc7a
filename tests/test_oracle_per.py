"""PER oracle self-consistency (CPU): tree invariants and proportional sampling."""
import numpy as np

from oracle.per_oracle import OraclePER


def test_tree_is_sum_of_children_and_total():
    per = OraclePER(1000, seed=1)
    rng = np.random.default_rng(0)
    per.fill(rng.random(1000).astype(np.float32))
    t = per.tree
    for node in (1, 2, 3, 17, 511, 1023):
        assert t[node] == np.float32(t[2 * node] + t[2 * node + 1])
    assert abs(float(per.total()) - float(t[per.L:].astype(np.float64).sum())) < 1e-2
    idx = rng.integers(0, 1000, 64)
    per.update(idx, rng.random(64).astype(np.float32))
    ref = OraclePER(1000, seed=1)
    ref.fill(t[per.L:per.L + 1000])
    assert np.array_equal(ref.tree, per.tree)           # incremental update == full rebuild


def test_sampling_is_proportional_and_in_support():
    per = OraclePER(64, seed=3)
    p = np.zeros(64, np.float32)
    p[[3, 10, 40]] = [1.0, 3.0, 6.0]
    per.fill(p)
    counts = np.zeros(64)
    for step in range(200):
        idx, pr = per.sample(step, 100)
        assert np.all(np.isin(idx, [3, 10, 40])) and np.array_equal(pr, p[idx])
        counts += np.bincount(idx, minlength=64)
    frac = counts / counts.sum()
    assert abs(frac[3] - 0.1) < 0.01 and abs(frac[10] - 0.3) < 0.01 and abs(frac[40] - 0.6) < 0.01
