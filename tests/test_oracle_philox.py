"""Philox4x32-10 oracle: Random123 known-answer vectors and the index derivation."""
import numpy as np

from oracle.philox import KAT_VECTORS, mulhi64, philox4x32_10, sample_indices


def test_random123_known_answers():
    for ctr, key, expect in KAT_VECTORS:
        out = philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in out) == expect


def test_vectorised_matches_scalar():
    ctr0 = np.arange(50, dtype=np.uint64)
    vec = philox4x32_10(ctr0, 3, 0, 9, 0xDEADBEEF, 0x1234)
    for i in range(50):
        one = philox4x32_10(i, 3, 0, 9, 0xDEADBEEF, 0x1234)
        assert all(int(vec[j][i]) == int(one[j]) for j in range(4))


def test_mulhi64_exact():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2**63, 200, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 200, dtype=np.uint64)
    for n in (1, 2, 3, 321, 100000, 1000000, 16 * 2**20, 2**31 - 1):
        got = mulhi64(x, n)
        want = np.array([(int(v) * n) >> 64 for v in x], dtype=np.int64)
        assert np.array_equal(got, want)


def test_indices_in_range_and_reproducible():
    for size in (1, 7, 321, 10**6):
        idx = sample_indices(seed=42, agent=3, step=11, batch_size=4096, size=size)
        assert idx.dtype == np.int64 and idx.min() >= 0 and idx.max() < size
        assert np.array_equal(idx, sample_indices(42, 3, 11, 4096, size))
    a = sample_indices(1, 0, 0, 64, 1000)
    assert not np.array_equal(a, sample_indices(1, 0, 1, 64, 1000))      # step changes the draw
    assert not np.array_equal(a, sample_indices(1, 1, 0, 64, 1000))      # agent changes the draw
    assert not np.array_equal(a, sample_indices(2, 0, 0, 64, 1000))      # seed changes the draw
    # prefix property: the first B draws of a step do not depend on B
    assert np.array_equal(sample_indices(1, 0, 0, 128, 1000)[:64], a)


def test_indices_roughly_uniform():
    idx = sample_indices(5, 0, 0, 200000, 10)
    counts = np.bincount(idx, minlength=10)
    assert np.all(np.abs(counts - 20000) < 5 * np.sqrt(20000))
