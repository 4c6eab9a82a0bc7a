"""Host-side mirror of the reference interface: specs, tree<->flat layout, sweep surface, sharding,
checkpoint layout.  No GPU."""
import inspect
import os
import pickle

import numpy as np
import pytest

import dqn_b200
from conftest import golden_tree
from oracle import dqn_oracle as O


def test_flat_layout_roundtrip_and_order(golden):
    params = golden_tree(golden["ref_checkpoint"], "params")
    flat = dqn_b200.flatten_tree(params, 9, 4)
    assert flat.dtype == np.float32 and flat.shape == (2757,) == (dqn_b200.param_count(9, 4),)
    # order: W1 b1 W2 b2 Wv bv Wa ba (include/dqn_b200.h)
    assert np.array_equal(flat[:288], params["model/~/linear"]["w"].reshape(-1))
    assert np.array_equal(flat[288 + 32:288 + 32 + 2048], params["model/~/linear_1"]["w"].reshape(-1))
    assert np.array_equal(flat[-4 - 256:-4], params["model/~/linear_3"]["w"].reshape(-1))
    back = dqn_b200.unflatten_tree(flat, 9, 4)
    for m in O.MODULES:
        for k in ("w", "b"):
            assert np.array_equal(back[m][k], params[m][k])
    with pytest.raises(ValueError):
        dqn_b200.flatten_tree(params, 8, 4)
    assert dqn_b200.param_count(8, 4) == 2725


def test_model_init_matches_haiku_defaults():
    tree = dqn_b200.Model(4).init(123, np.zeros((1, 9), np.float32))
    assert list(tree) == list(O.MODULES)
    for (fi, fo), m in zip([(9, 32), (32, 64), (64, 1), (64, 4)], O.MODULES):
        w, b = tree[m]["w"], tree[m]["b"]
        assert w.shape == (fi, fo) and w.dtype == np.float32 and not b.any()
        assert np.abs(w).max() <= 2.0 / np.sqrt(fi) + 1e-6          # truncated at 2 sigma
    with pytest.raises(ValueError):
        dqn_b200.Model(4, hidden=(64, 64))


def test_optimizer_specs_match_the_two_scripts():
    w = dqn_b200.adamw(2e-4)        # Test/lunar_lander.py:48
    a = dqn_b200.adam(1e-4)         # Test/lunar_lander_hyper_params.py:41
    assert (w.kind, w.weight_decay, w.b1, w.b2, w.eps, w.eps_root) == ("adamw", 1e-4, 0.9, 0.999, 1e-8, 0.0)
    assert (a.kind, a.weight_decay) == ("adam", 0.0)
    params = dqn_b200.Model(4).init(0, np.zeros((1, 9), np.float32))
    st = w.init(params)
    assert len(st) == 3 and type(st[0]).__name__ == "ScaleByAdamState" and int(st[0].count) == 0
    assert len(a.init(params)) == 2
    assert list(st[0].mu["model/~/linear"]) == ["b", "w"]           # key order of the reference pickle


def test_agent_constructor_keeps_reference_keywords():
    want = ["network", "params", "optimizer", "opt_state", "env", "buffer_size", "obs_shape", "ac_shape", "gamma",
            "epsilon", "epsilon_decay_rate", "min_epsilon", "max_episodes", "max_steps", "training_start",
            "batch_size", "train_frequency", "back_up_frequency", "replace_frequency", "reward_to_reach",
            "num_actions", "saving_directory", "monitoring", "verbose"]          # q_agent.py:61-86
    got = [p for p in inspect.signature(dqn_b200.Agent.__init__).parameters if p != "self"]
    assert got[:len(want)] == want
    for name in ("_policy", "_step", "_update_target_model", "_run_episode", "training", "evaluate",
                 "_update_epsilon", "_update_reward_history", "_average_reward"):
        assert hasattr(dqn_b200.Agent, name)
    pa = inspect.signature(dqn_b200.ParamAgent.__init__).parameters
    assert pa["gamma"].default == 0. and pa["batch_size"].default == 0     # hyperparameter_optimization.py:33-39
    inj = [p for p in inspect.signature(dqn_b200.ParamAgent.inject).parameters if p != "self"]
    assert inj == ["gamma", "epsilon", "epsilon_decay_rate", "min_epsilon", "replace_frequency", "batch_size",
                   "train_frequency"]                                       # :76-84


def test_sweep_bounds_and_points():
    b = dqn_b200.SWEEP_BOUNDS                                               # hyperparameter_optimization.py:115-123
    assert b["batch_size"] == [38, 70] and b["gamma"] == [0.9, 0.999] and b["train_frequency"] == [2, 15]
    pts = dqn_b200.sweep_hparams(64, seed=1000)
    assert all(38 <= p["batch_size"] <= 69 and isinstance(p["batch_size"], int) for p in pts)
    assert all(0.9 <= p["gamma"] <= 0.999 for p in pts)
    assert pts == dqn_b200.sweep_hparams(64, seed=1000)                       # deterministic per global agent id
    assert dqn_b200.sweep_hparams(8, seed=1000) == pts[:8]


@pytest.mark.parametrize("n,world", [(1024, 1), (1024, 2), (1024, 8), (10, 4), (3, 8), (1, 2)])
def test_shard_range_partitions_agents(n, world):
    covered = []
    sizes = []
    for r in range(world):
        b, e = dqn_b200.shard_range(n, r, world)
        covered += list(range(b, e))
        sizes.append(e - b)
    assert covered == list(range(n)) and max(sizes) - min(sizes) <= 1


def test_checkpoint_roundtrip_in_reference_layout(tmp_path, golden):
    params = golden_tree(golden["ref_checkpoint"], "params")
    opt = dqn_b200.adamw(2e-4)
    st = opt.init(params)
    d = str(tmp_path / "lunar_lander")
    dqn_b200.generate_saving(d)(params, st)
    assert sorted(os.listdir(d)) == ["opt_state.pickle", "params.pickle"]      # General/Base/utils.py:25-28
    p2, s2 = dqn_b200.generate_loading(d)()
    for m in O.MODULES:
        for k in ("w", "b"):
            assert np.array_equal(p2[m][k], params[m][k])
    assert int(s2[0].count) == 0 and len(s2) == 3


def test_checkpoint_loads_with_bare_pickle(tmp_path, golden):
    """The reference's generate_loading (General/Base/utils.py:31-38) is a bare pickle.load in a process that has optax
    and knows nothing of this package: the files written here must name optax's state classes, not this package's."""
    import subprocess
    import sys
    import textwrap
    params = golden_tree(golden["ref_checkpoint"], "params")
    opt = dqn_b200.adamw(2e-4)
    d = str(tmp_path / "ckpt")
    dqn_b200.generate_saving(d)(params, opt.init(params))
    fake = tmp_path / "site"                       # the two optax modules the pickle names, as optax 0.1.x defines the classes
    (fake / "optax" / "_src").mkdir(parents=True)
    (fake / "optax" / "__init__.py").write_text("")
    (fake / "optax" / "_src" / "__init__.py").write_text("")
    (fake / "optax" / "_src" / "transform.py").write_text(
        "from typing import Any, NamedTuple\nclass ScaleByAdamState(NamedTuple):\n    count: Any\n    mu: Any\n    nu: Any\n")
    (fake / "optax" / "_src" / "base.py").write_text("from typing import NamedTuple\nclass EmptyState(NamedTuple):\n    pass\n")
    code = textwrap.dedent("""
        import pickle, sys
        sys.path.insert(0, %r)
        params = pickle.load(open(%r, "rb"))
        st = pickle.load(open(%r, "rb"))
        assert "deep_q_learning_b200" not in sys.modules and "dqn_b200" not in sys.modules
        assert type(st[0]).__module__ == "optax._src.transform" and type(st[1]).__module__ == "optax._src.base"
        assert int(st[0].count) == 0 and len(st) == 3 and st[0].mu["model/~/linear"]["w"].shape == (9, 32)
        assert params["model/~/linear_3"]["w"].shape == (64, 4) and params["model/~/linear"]["w"].dtype.name == "float32"
        print("ok")
    """) % (str(fake), os.path.join(d, "params.pickle"), os.path.join(d, "opt_state.pickle"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stderr


def test_add_accepts_strided_observation_with_or_without_hoststage():
    """ReplayBuffer.add must not depend on whether the optional C staging helper is built: a non-contiguous observation
    makes the helper refuse the buffer (BufferError / ValueError), and the numpy path takes over."""
    hs = dqn_b200.pkg.replay._hoststage
    if hs is None:
        pytest.skip("_hoststage not built")
    D, cap = 9, 4
    a = [np.zeros((cap, D), np.float32), np.zeros(cap, np.int64), np.zeros(cap, np.float32), np.zeros((cap, D), np.float32), np.zeros(cap, np.bool_)]
    st = hs.new(*[x.ctypes.data for x in a], D, cap)
    strided = np.arange(2 * D, dtype=np.float32)[::2]
    with pytest.raises((TypeError, BufferError, ValueError)):
        hs.put(st, 0, strided, 0, 0.0, np.zeros(D, np.float32), False)


def test_loader_reads_the_reference_pickles_when_present(golden):
    ref = "/root/reference/Test/lunar_lander"
    if not os.path.exists(ref):
        pytest.skip("reference tree only exists in the build container")
    params, opt_state = dqn_b200.generate_loading(ref)()
    gold = golden_tree(golden["ref_checkpoint"], "params")
    for m in O.MODULES:
        for k in ("w", "b"):
            assert np.array_equal(params[m][k], gold[m][k])
    assert int(opt_state[0].count) == 0 and len(opt_state) == 3


def test_sample_batch_rejects_host_arrays():
    z = np.zeros((4, 9), np.float32)
    with pytest.raises(TypeError):
        dqn_b200.sample_batch(4, z, np.zeros(4, np.int64), np.zeros(4, np.float32), z, np.zeros(4, bool), 2)


def test_hoststage_matches_numpy_stores():
    """csrc/hoststage.c (optional CPython accelerator of ReplayBuffer.add's host staging) stores exactly what the numpy
    assignments store, for the argument types the reference passes (float32 rows, python / numpy scalars)."""
    hs = dqn_b200.pkg.replay._hoststage
    if hs is None:
        pytest.skip("_hoststage not built")
    D, cap = 9, 16
    a = [np.zeros((cap, D), np.float32), np.zeros(cap, np.int64), np.zeros(cap, np.float32), np.zeros((cap, D), np.float32), np.zeros(cap, np.bool_)]
    b = [x.copy() for x in a]
    st = hs.new(*[x.ctypes.data for x in a], D, cap)
    rng = np.random.default_rng(0)
    for i in range(cap):
        s, o = rng.standard_normal(D).astype(np.float32), rng.standard_normal((1, D)).astype(np.float32)[0]
        act = [3, np.int64(2), np.int32(1), -1][i % 4]
        rew = [1.25, np.float32(-0.1), np.float64(1e-3), 7][i % 4]
        done = [True, False, np.bool_(True), 0][i % 4]
        hs.put(st, i, s, act, rew, o, done)
        b[0][i], b[1][i], b[2][i], b[3][i], b[4][i] = s, act, rew, o, done
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    with pytest.raises(TypeError):
        hs.put(st, 0, np.zeros(D, np.float64), 0, 0.0, np.zeros(D, np.float32), False)       # replay.py falls back to numpy here
    with pytest.raises(IndexError):
        hs.put(st, cap, np.zeros(D, np.float32), 0, 0.0, np.zeros(D, np.float32), False)
