"""GPU parity of the replay ring (through the C ABI) against the reference's own outputs (golden) and
the oracle: stored contents, gathered minibatches and Philox indices are BIT-EXACT."""
import numpy as np
import pytest

import dqn_b200
from oracle.philox import sample_indices
from oracle.replay_oracle import OracleReplay, gather, synthetic_transitions

pytestmark = pytest.mark.gpu
FIELDS = ("states", "actions", "rewards", "observations", "dones")


def assert_arrays_equal(got, ref, what=""):
    for name, g, r in zip(FIELDS, got, ref):
        g = np.asarray(g)
        assert g.dtype == r.dtype, (what, name, g.dtype, r.dtype)
        assert g.shape == r.shape, (what, name, g.shape, r.shape)
        assert np.array_equal(g.view(np.uint8), r.view(np.uint8)), (what, name)


def test_add_stream_bit_exact_vs_reference(golden):
    g = golden["replay_ref_stream"]
    N, D = int(g["N"]), int(g["D"])
    buf = dqn_b200.ReplayBuffer(N, (N, D), (N,))
    marks = set(int(m) for m in g["marks"])
    for i in range(len(g["in_actions"])):
        buf.add(g["in_states"][i], int(g["in_actions"][i]), float(g["in_rewards"][i]),
                g["in_observations"][i], bool(g["in_dones"][i]))
        if i + 1 in marks:
            k = f"after{i + 1}"
            got = [np.asarray(getattr(buf, n)) for n in FIELDS]
            assert_arrays_equal(got, [g[f"{k}_{n}"] for n in FIELDS], k)
            assert buf.size == int(g[f"{k}_size"]) and buf._counter == int(g[f"{k}_counter"])
            assert buf._engine.buffer_state(0) == (int(g[f"{k}_size"]), int(g[f"{k}_counter"]))


def test_vectorised_store_equals_scalar_adds(golden):
    g = golden["replay_ref_stream"]
    N, D = int(g["N"]), int(g["D"])
    ins = [g["in_states"], g["in_actions"], g["in_rewards"].astype(np.float32), g["in_observations"], g["in_dones"]]
    for chunks in ([100], [1, 19, 17, 1, 26, 36], [37, 37, 26], [99, 1]):
        buf = dqn_b200.ReplayBuffer(N, (N, D), (N,))
        o = 0
        for c in chunks:
            buf.add_many(*[x[o:o + c] for x in ins])
            o += c
        got = [np.asarray(getattr(buf, n)) for n in FIELDS]
        assert_arrays_equal(got, [g[f"after100_{n}"] for n in FIELDS], str(chunks))
        assert buf.size == N and buf._counter == 100


def test_sample_batch_bit_exact_vs_reference(golden):
    g = golden["replay_ref_sample"]
    N, D, filled = int(g["N"]), int(g["D"]), int(g["filled"])
    buf = dqn_b200.ReplayBuffer(N, (N, D), (N,))
    buf.add_many(*[g[n][:filled] for n in FIELDS])
    assert buf.size == filled
    for c in range(int(g["n_cases"])):
        B = int(g[f"case{c}_B"])
        out = dqn_b200.sample_batch(buf.size, buf.states, buf.actions, buf.rewards, buf.observations, buf.dones,
                                    B, indices=g[f"case{c}_idx"])
        assert_arrays_equal(out, [g[f"case{c}_{n}"] for n in FIELDS], f"case{c}")


@pytest.mark.parametrize("D,N,n", [(8, 1000, 2500), (9, 777, 777), (16, 300, 1000), (1, 64, 10), (3, 5, 23)])
def test_store_wraparound_and_obs_dims_vs_oracle(D, N, n):
    rng = np.random.default_rng(D * 1000 + N)
    s, a, r, s2, d = synthetic_transitions(rng, n, D, 4, done_p=0.3)
    a = a.astype(np.int64) + rng.integers(-3, 3, n) * (1 << 40)     # full int64 range survives the ring
    ora = OracleReplay(N, (N, D), (N,))
    buf = dqn_b200.ReplayBuffer(N, (N, D), (N,))
    o = 0
    while o < n:
        c = int(rng.integers(1, max(2, n // 3)))
        ora.add_many(s[o:o + c], a[o:o + c], r[o:o + c], s2[o:o + c], d[o:o + c])
        buf.add_many(s[o:o + c], a[o:o + c], r[o:o + c], s2[o:o + c], d[o:o + c])
        o += c
    assert_arrays_equal([np.asarray(getattr(buf, f)) for f in FIELDS], ora.arrays())
    assert buf.size == ora.size and buf._counter == ora.counter
    idx = rng.integers(0, ora.size, 257)
    assert_arrays_equal(buf.sample(257, indices=idx), gather(idx, *ora.arrays()))


def test_special_float_payloads_survive():
    N, D = 16, 9
    buf = dqn_b200.ReplayBuffer(N, (N, D), (N,))
    s = np.zeros((4, D), np.float32)
    s[0, :4] = [np.nan, np.inf, -np.inf, -0.0]
    s[1] = np.float32(1e-45)                       # subnormal
    s.view(np.uint32)[2, 0] = 0x7FC12345           # NaN payload
    r = np.array([np.nan, -0.0, 3.4e38, 1e-45], np.float32)
    buf.add_many(s, np.arange(4), r, s[::-1].copy(), np.array([1, 0, 1, 0], bool))
    got = [np.asarray(getattr(buf, f)) for f in FIELDS]
    assert np.array_equal(got[0][:4].view(np.uint32), s.view(np.uint32))
    assert np.array_equal(got[2][:4].view(np.uint32), r.view(np.uint32))
    assert np.array_equal(got[3][:4].view(np.uint32), s[::-1].view(np.uint32))


@pytest.mark.parametrize("size,B", [(1, 64), (321, 70), (100000, 4096), (999983, 64)])
def test_philox_indices_bit_exact_vs_oracle(size, B):
    N, D = size + 5, 8
    eng = dqn_b200.DqnEngine(D, 4, N, 64, 0.99, dqn_b200.adam(1e-4), seed=0xC0FFEE1234, agent_id_base=3)
    rng = np.random.default_rng(0)
    eng.store(*synthetic_transitions(rng, size, D))
    for step in (0, 1, 17, 2**33 + 5):
        got = eng.sample_indices(step, B)
        want = sample_indices(0xC0FFEE1234, 3, step, B, size)
        assert got.dtype == np.int64 and np.array_equal(got, want), (size, B, step)


def test_philox_sample_batch_matches_explicit_gather():
    N, D = 5000, 8
    eng = dqn_b200.DqnEngine(D, 4, N, 64, 0.99, dqn_b200.adam(1e-4), seed=99)
    rng = np.random.default_rng(1)
    data = synthetic_transitions(rng, 3210, D)
    eng.store(*data)
    idx = sample_indices(99, 0, 12, 512, 3210)
    assert_arrays_equal(eng.sample_batch(512, step=12), gather(idx, *data))


def test_empty_ring_and_bad_indices_raise():
    eng = dqn_b200.DqnEngine(8, 4, 100, 8, 0.99, dqn_b200.adam(1e-4))
    with pytest.raises(dqn_b200.DqnError):
        eng.sample_batch(8)                          # randint(0, 0) raises in the reference too
    with pytest.raises(dqn_b200.DqnError):
        eng.train_steps(1)
    eng.store(*synthetic_transitions(np.random.default_rng(0), 10, 8))
    with pytest.raises(dqn_b200.DqnError):
        eng.train_steps(1, indices=np.full(8, 10))   # slot 10 is beyond size
    s = eng.sample_batch(0)
    assert s[0].shape == (0, 8)


def test_full_size_ring_properties():
    """BASELINE config 2 size (1M slots, D=8): store -> export round trip, checksum of checksums, and a
    large gather whose result equals the exported arrays at the drawn slots."""
    N, D = 1_000_000, 8
    eng = dqn_b200.DqnEngine(D, 4, N, 64, 0.99, dqn_b200.adam(1e-4), seed=5)
    rng = np.random.default_rng(2)
    data = synthetic_transitions(rng, N + 12345, D)          # wraps by 12345
    for o in range(0, N + 12345, 250_000):
        eng.store(*[x[o:o + 250_000] for x in data])
    exp = eng.buffer_export()
    pos = (np.arange(N + 12345) % N)[-N:]
    want = [np.empty_like(e) for e in exp]
    for w, x in zip(want, data):
        w[pos] = x[-N:]
    assert_arrays_equal(exp, want)
    assert int(exp[0].view(np.uint32).sum(dtype=np.uint64)) == int(want[0].view(np.uint32).sum(dtype=np.uint64))
    idx = eng.sample_indices(3, 65536 // 2)
    assert idx.min() >= 0 and idx.max() < N and len(np.unique(idx)) > 30000      # with replacement, spread out
    got = eng.sample_batch(65536 // 2, indices=idx)
    assert_arrays_equal(got, [w[idx] for w in want])
