"""GPU parity of the fused train step (through the C ABI) against the oracle on identical weights,
transitions and indices.  Bar (north_star): indices / greedy actions bit-exact; Q-values, TD targets,
loss, gradients and updated parameters within 1e-5 relative in fp32 (assert_close adds an absolute
floor of 1e-6 x the tensor's largest magnitude for entries that are ~0 by cancellation)."""
import numpy as np
import pytest

import dqn_b200
from conftest import assert_close, golden_tree
from oracle import dqn_oracle as O
from oracle.agent_oracle import OracleAgent
from oracle.philox import sample_indices
from oracle.replay_oracle import synthetic_transitions

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["cta", "cluster", "cta_tc"])
def step_kernel(request, monkeypatch):
    """Every test of this module runs against both train-step kernels: one CTA per agent (train_fused.cu) and
    one agent over a 4-CTA cluster (train_cluster.cu)."""
    monkeypatch.setenv("DQN_B200_STEP_KERNEL", request.param)
    return request.param


def make_pair(D=9, A=4, B=64, N=2000, n_fill=1500, kind="adamw", lr=2e-4, gamma=0.99, seed=0, params=None,
              done_p=0.2, perturb_target=True):
    rng = np.random.default_rng(seed)
    if params is None:
        params = O.init_params(rng, D, A, bias_std=0.05)
    else:                                   # fixture theta_0 has zero biases: randomise them
        params = O.tree_copy(params)
        for m in O.MODULES:
            params[m]["b"] = (0.05 * rng.standard_normal(params[m]["b"].shape)).astype(np.float32)
    target = O.tree_map(lambda x: (x + 0.02 * rng.standard_normal(x.shape)).astype(np.float32), params) \
        if perturb_target else O.tree_copy(params)
    opt = dqn_b200.adamw(lr) if kind == "adamw" else dqn_b200.adam(lr)
    eng = dqn_b200.DqnEngine(D, A, N, B, gamma, opt, seed=seed + 100)
    eng.set_params(params, 0, 0)
    eng.set_params(target, 0, 1)
    ora = OracleAgent(params, O.init_opt_state(params), O.OptSpec(kind, lr), N, D, gamma, B, seed=seed + 100)
    ora.target_params = O.tree_copy(target)
    data = synthetic_transitions(rng, n_fill, D, A, done_p=done_p)
    eng.store(*data)
    ora.replay.add_many(*data)
    return eng, ora, rng


def compare_step(eng, ora, indices=None, what=""):
    ref = ora.step(indices)
    got = eng.train_step_debug(indices=indices)
    assert np.array_equal(got["indices"], ref["indices"]), what + " indices"
    assert np.array_equal(got["max_actions"], ref["max_actions"]), what + " argmax"
    for k in ("q", "next_q", "next_q_tm", "targets"):
        assert_close(got[k], ref[k], what=f"{what} {k}")
    assert abs(float(got["loss"]) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"])) + 1e-9, what + " loss"
    for m in O.MODULES:
        for k in ("w", "b"):
            assert_close(got["grads"][m][k], ref["grads"][m][k], what=f"{what} grad {m}/{k}")
    compare_state(eng, ora, what)
    return got, ref


def compare_state(eng, ora, what="", rtol=1e-5):
    p = eng.get_params(0, 0)
    cnt, mu, nu = eng.get_opt_state(0)
    assert int(cnt) == int(ora.opt_state["count"]), what + " adam count"
    for m in O.MODULES:
        for k in ("w", "b"):
            assert_close(p[m][k], ora.params[m][k], rtol=rtol, what=f"{what} param {m}/{k}")
            assert_close(mu[m][k], ora.opt_state["mu"][m][k], rtol=rtol, what=f"{what} mu {m}/{k}")
            assert_close(nu[m][k], ora.opt_state["nu"][m][k], rtol=rtol, what=f"{what} nu {m}/{k}")


def test_lunar_lander_config_from_reference_checkpoint(golden):
    """Config 1 (Test/lunar_lander.py): D=9, A=4, B=64, gamma .99, AdamW(2e-4), theta_0 = the shipped pickle."""
    theta0 = golden_tree(golden["ref_checkpoint"], "params")
    eng, ora, rng = make_pair(params=theta0)
    for step in range(5):
        compare_step(eng, ora, what=f"step{step}")          # Philox indices, drawn on both sides
    t = eng.get_params(0, 1)
    for m in O.MODULES:                                     # target network untouched by training
        assert np.array_equal(t[m]["w"], ora.target_params[m]["w"])


@pytest.mark.parametrize("D,A,B,kind,gamma", [
    (8, 4, 64, "adam", 0.9028), (9, 4, 38, "adam", 0.0), (9, 4, 70, "adamw", 0.999), (8, 4, 1, "adam", 0.95),
    (8, 4, 128, "adamw", 0.99), (16, 2, 33, "adam", 0.9), (1, 7, 65, "adamw", 0.5), (4, 3, 200, "adam", 0.99),
    (16, 5, 80, "adam", 0.97), (16, 2, 72, "adamw", 0.9), (8, 4, 81, "adam", 0.99),     # 65..80: one tile with tail rows; 81: two tiles
])
def test_shapes_batches_optimisers(D, A, B, kind, gamma):
    eng, ora, rng = make_pair(D=D, A=A, B=B, kind=kind, lr=1e-3, gamma=gamma, seed=D * 7 + B)
    for step in range(3):
        compare_step(eng, ora, what=f"D{D} A{A} B{B} step{step}")


def test_edge_sizes():
    """Largest batch the ABI accepts (16 tiles of 64 rows), a one-slot ring, a ring smaller than the batch."""
    eng, ora, rng = make_pair(D=8, A=4, B=1024, N=3000, n_fill=3000, lr=1e-3, seed=31)
    compare_step(eng, ora, what="B=1024")
    compare_step(eng, ora, what="B=1024 step 2")
    eng, ora, rng = make_pair(D=9, A=4, B=16, N=1, n_fill=1, seed=32)            # every sample is the same transition
    compare_step(eng, ora, what="N=1")
    eng, ora, rng = make_pair(D=9, A=4, B=64, N=10, n_fill=25, seed=33)           # wrapped 2.5 times, size 10 < B
    compare_step(eng, ora, what="N=10 < B")
    with pytest.raises(dqn_b200.DqnError):
        dqn_b200.DqnEngine(9, 4, 100, 1025, 0.99, dqn_b200.adam(1e-3))             # B > DQN_MAX_BATCH
    empty = dqn_b200.DqnEngine(9, 4, 100, 8, 0.99, dqn_b200.adam(1e-3))
    with pytest.raises(dqn_b200.DqnError):
        empty.train_steps(1)                                                       # randint(0, 0) raises in the reference too


def test_explicit_indices_and_duplicates():
    eng, ora, rng = make_pair(B=64)
    idx = rng.integers(0, 1500, 64)
    idx[:8] = idx[8]                                        # with-replacement duplicates
    compare_step(eng, ora, indices=idx, what="explicit")
    compare_step(eng, ora, indices=np.zeros(64, np.int64), what="all-same-slot")


def test_terminal_rows_follow_reference_quirk_F5():
    eng, ora, rng = make_pair(done_p=1.0)
    got, ref = compare_step(eng, ora, what="all-terminal")
    batch_r = ora.replay.rewards[ref["indices"]]
    rows = np.arange(64)
    a = ora.replay.actions[ref["indices"]]
    np.testing.assert_allclose(got["targets"][rows, a] - got["q"][rows, a], batch_r, rtol=1e-5, atol=1e-6)


def test_hundred_step_drift_and_hard_sync():
    """100 consecutive steps (hard sync every 25) stay within tolerance of the oracle."""
    eng, ora, rng = make_pair(seed=3, lr=1e-3, perturb_target=False)
    for step in range(100):
        ora.step()
        eng.train_steps(1)
        if step % 25 == 24:
            ora.update_target_model()
            eng.sync_target()
            t = eng.get_params(0, 1)
            o = eng.get_params(0, 0)
            for m in O.MODULES:
                assert np.array_equal(t[m]["w"], o[m]["w"]) and np.array_equal(t[m]["b"], o[m]["b"])
    compare_state(eng, ora, "after 100 steps", rtol=2e-4)     # fp32 summation-order drift accumulates
    losses = eng.losses(100)
    assert_close(losses[-1], np.float32(ora.last["loss"]), rtol=1e-3, what="loss[99]")
    assert eng.train_step_count() == 100


def test_fused_k_steps_equal_k_single_launches_bitwise():
    """K steps in one persistent launch == K launches of one step (same Philox stream): bit-identical."""
    eng1, _, _ = make_pair(seed=4, lr=1e-3)
    eng2, _, _ = make_pair(seed=4, lr=1e-3)
    eng3, _, _ = make_pair(seed=4, lr=1e-3)
    for _ in range(37):
        eng1.train_steps(1)
    eng2.train_steps(37)
    eng3.train_steps(20)
    eng3.train_steps(17)
    a, b, c = eng1.get_params_flat(), eng2.get_params_flat(), eng3.get_params_flat()
    assert np.array_equal(a, b) and np.array_equal(a, c)
    for e in (eng2, eng3):
        c1, m1, v1 = eng1.get_opt_state()
        c2, m2, v2 = e.get_opt_state()
        assert c1 == c2 == 37
        for m in O.MODULES:
            assert np.array_equal(m1[m]["w"], m2[m]["w"]) and np.array_equal(v1[m]["w"], v2[m]["w"])
    assert np.array_equal(eng1.losses(37), eng2.losses(37))


def test_store_and_train_in_one_launch_equals_store_then_train():
    """dqn_store_train_step: n add()s + one _step() as ONE launch (transitions in the kernel-parameter buffer) ==
    dqn_store followed by dqn_train_step, bit for bit -- ring contents (incl. wrap-around), parameters, losses."""
    _lib = dqn_b200.pkg._lib
    D, A, N, B = 9, 4, 50, 32
    rng = np.random.default_rng(12)
    theta = O.init_params(rng, D, A, bias_std=0.05)
    engs = [dqn_b200.DqnEngine(D, A, N, B, 0.99, dqn_b200.adamw(1e-3), seed=77) for _ in range(2)]
    ora = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adamw", 1e-3), N, D, 0.99, B, seed=77)
    for e in engs:
        e.set_params(theta, 0, 0)
        e.set_params(theta, 0, 1)
    loss = np.zeros(1, np.float32)
    for it, n in enumerate([3, 1, 16, 4, 4, 7, 16, 16, 2, 17, 4, 0, 5]):      # 17 > DQN_MAX_INLINE_STORE -> fallback path
        s, a, r, s2, d = synthetic_transitions(rng, max(n, 1), D, A, done_p=0.2)
        s, a, r, s2, d = s[:n], a[:n], r[:n], s2[:n], d[:n]
        if n:
            engs[0].store(s, a, r, s2, d)
            ora.replay.add_many(s, a, r, s2, d)
        engs[0].train_steps(1)
        ora.step()
        d8 = np.ascontiguousarray(d, dtype=np.bool_)
        _lib.check(engs[1].lib.dqn_store_train_step(engs[1].h, 0, n, _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(s2),
                                                    _lib.ptr(d8), 1, _lib.ptr(loss)))
        assert loss[0] == engs[0].last_loss(), f"iteration {it}"
        assert engs[1].buffer_state() == engs[0].buffer_state()
    assert np.array_equal(engs[0].get_params_flat(), engs[1].get_params_flat())
    for x, y, z in zip(engs[0].buffer_export(), engs[1].buffer_export(),
                       (ora.replay.states, ora.replay.actions, ora.replay.rewards, ora.replay.observations, ora.replay.dones)):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    compare_state(engs[1], ora, "store+train fused", rtol=2e-5)


def test_long_runs_are_bitwise_reproducible():
    """4000 steps, once as 8 launches of 500 and once as 40 launches of 100, on two handles: bit-identical parameters,
    moments and losses.  (The cluster kernel exchanges gradients and weights through distributed shared memory with
    relaxed cluster barriers; a missed ordering would show up here as run-to-run divergence.)"""
    a, _, _ = make_pair(seed=11, lr=1e-3, N=5000, n_fill=5000)
    b, _, _ = make_pair(seed=11, lr=1e-3, N=5000, n_fill=5000)
    for _ in range(8):
        a.train_steps(500)
    for _ in range(40):
        b.train_steps(100)
    assert np.array_equal(a.get_params_flat(), b.get_params_flat())
    (ca, ma, va), (cb, mb, vb) = a.get_opt_state(), b.get_opt_state()
    assert ca == cb == 4000
    for m in O.MODULES:
        for k in ("w", "b"):
            assert np.array_equal(ma[m][k], mb[m][k]) and np.array_equal(va[m][k], vb[m][k])
    assert np.array_equal(a.losses(4000), b.losses(4000))
    assert np.all(np.isfinite(a.get_params_flat()))


def test_greedy_actions_bit_exact():
    eng, ora, rng = make_pair(seed=9)
    states = rng.standard_normal((2000, 9)).astype(np.float32)
    q = O.forward(ora.params, states)
    srt = np.sort(q, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-5               # exclude numerical near-ties (none expected)
    assert clear.mean() > 0.99
    want = np.argmax(q, axis=1)
    got = np.array([eng.act(s) for s in states[:300]])
    assert np.array_equal(got[clear[:300]], want[:300][clear[:300]])
    zero = O.tree_zeros_like(ora.params)
    eng.set_params(zero, 0, 0)
    assert eng.act(states[0]) == 0                          # all-equal Q -> first index, like numpy/jnp argmax


def test_agent_dropin_matches_oracle_agent(golden):
    """The reference-facing Agent object: constructor kwargs of Test/lunar_lander.py, add() per transition,
    _step(), _update_target_model(), _policy() greedy branch."""
    import asyncio
    theta0 = golden_tree(golden["ref_checkpoint"], "params")
    opt = dqn_b200.adamw(0.0002)
    N, D, B = 1000, 9, 64
    agent = dqn_b200.Agent(network=dqn_b200.Model(4), params=theta0, optimizer=opt, opt_state=opt.init(theta0),
                           env=None, buffer_size=N, obs_shape=(N, D), ac_shape=(N,), gamma=0.99, epsilon=0.0,
                           epsilon_decay_rate=0.99, min_epsilon=0.0, max_episodes=10, max_steps=1500,
                           training_start=250, batch_size=B, train_frequency=4, back_up_frequency=50,
                           replace_frequency=20, reward_to_reach=230.0, num_actions=4,
                           saving_directory="/tmp/dqn_b200_test_agent", monitoring=False, seed=21)
    ora = OracleAgent(theta0, O.init_opt_state(theta0), O.OptSpec("adamw", 2e-4), N, D, 0.99, B, seed=21)
    rng = np.random.default_rng(8)
    s, a, r, s2, d = synthetic_transitions(rng, 300, D, 4, done_p=0.2)
    for i in range(300):
        agent._replay_buffer.add(s[i], int(a[i]), float(r[i]), s2[i], bool(d[i]))
        ora.add(s[i], int(a[i]), float(r[i]), s2[i], bool(d[i]))
        if i >= 250 and i % 4 == 0:
            agent._step()
            ora.step()
        if i == 280:
            asyncio.run(agent._update_target_model())
            ora.update_target_model()
    p = agent._params
    for m in O.MODULES:
        for k in ("w", "b"):
            assert_close(p[m][k], ora.params[m][k], what=f"agent param {m}/{k}")
    assert agent._replay_buffer.size == 300
    st = agent._opt_state
    assert int(st[0].count) == int(ora.opt_state["count"]) and len(st) == 3
    assert agent._policy(s[:1]) == ora.policy_greedy(s[:1])          # epsilon = 0 -> always greedy
    agent._epsilon = 2.0                                             # epsilon > 1 -> always random
    assert all(0 <= agent._policy(s[:1]) < 4 for _ in range(20))


def test_param_agent_inject_and_gamma_modes():
    theta = O.init_params(np.random.default_rng(0), 9, 4, bias_std=0.05)
    opt = dqn_b200.adam(1e-4)
    kw = dict(network=dqn_b200.Model(4), params=theta, optimizer=opt, opt_state=opt.init(theta), env=None,
              buffer_size=500, obs_shape=(500, 9), ac_shape=(500,), max_episodes=10, max_steps=1500,
              training_start=500, back_up_frequency=50, reward_to_reach=240.0, num_actions=4,
              saving_directory="/tmp/dqn_b200_test_pagent")
    data = synthetic_transitions(np.random.default_rng(1), 400, 9, 4, done_p=0.1)
    for mode, eff_gamma in (("frozen", 0.0), ("live", 0.9028)):
        ag = dqn_b200.ParamAgent(**kw, gamma_mode=mode, seed=5)
        ag.inject(0.9028, 0.979, 0.9873, 0.1469, 25, 52, 7)           # the optimum recorded in the sweep script
        assert ag._engine.get_hparams(0)["batch_size"] == 52 and ag._batch_size == 52
        assert abs(ag._engine.get_hparams(0)["gamma"] - eff_gamma) < 1e-7
        ag.max_episodes = 500
        assert ag.max_episodes == 500
        ag._replay_buffer.add_many(*data)
        ora = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-4), 500, 9, eff_gamma, 52, seed=5)
        ora.replay.add_many(*data)
        for _ in range(3):
            ag._step()
            ora.step()
        p = ag._params
        for m in O.MODULES:
            assert_close(p[m]["w"], ora.params[m]["w"], what=f"{mode} {m}")


def test_resume_state_continues_bitwise(tmp_path, golden):
    """save_resume_state / load_resume_state: a fresh agent restored from the file continues exactly like
    the one that kept running -- parameters, both networks, moments, ring (wrapped), Philox position, decay powers."""
    theta0 = golden_tree(golden["ref_checkpoint"], "params")

    def make():
        opt = dqn_b200.adamw(1e-3)
        return dqn_b200.Agent(network=dqn_b200.Model(4), params=theta0, optimizer=opt, opt_state=opt.init(theta0), env=None,
                              buffer_size=150, obs_shape=(150, 9), ac_shape=(150,), gamma=0.97, epsilon=0.7, epsilon_decay_rate=0.99,
                              min_epsilon=0.1, max_episodes=10, max_steps=100, training_start=10, batch_size=48, train_frequency=2,
                              back_up_frequency=50, replace_frequency=3, reward_to_reach=1e9, num_actions=4,
                              saving_directory=str(tmp_path), seed=19)
    rng = np.random.default_rng(6)
    s, a, r, s2, d = synthetic_transitions(rng, 400, 9, 4, done_p=0.2)

    def drive(agent, lo, hi):
        for i in range(lo, hi):
            agent._replay_buffer.add(s[i], int(a[i]), float(r[i]), s2[i], bool(d[i]))
            if i >= 10 and i % 2 == 0:
                agent._step()
            if i % 60 == 59:
                agent._sync_target()

    a1 = make()
    drive(a1, 0, 230)                                         # the 150-slot ring has wrapped
    a1._epsilon, a1._reward_history = 0.4321, [1.5, -2.0, 7.25]
    path = str(tmp_path / "resume.npz")
    a1.save_resume_state(path)
    a2 = make()
    a2.load_resume_state(path)
    assert a2._epsilon == 0.4321 and a2._reward_history == [1.5, -2.0, 7.25] and a2._replay_buffer.size == 150
    drive(a1, 230, 400)
    drive(a2, 230, 400)
    e1, e2 = a1._engine, a2._engine
    assert np.array_equal(e1.get_params_flat(0, 0), e2.get_params_flat(0, 0))
    assert np.array_equal(e1.get_params_flat(0, 1), e2.get_params_flat(0, 1))
    assert e1.get_counters() == e2.get_counters()
    for x, y in zip(e1.buffer_export(), e2.buffer_export()):
        assert np.array_equal(x, y)
    assert np.array_equal(e1.losses(50), e2.losses(50))


def test_adam_update_from_own_gradient():
    """The optimiser arithmetic in isolation: mu', nu', theta' recomputed on the host (oracle adam_update, fp32) FROM THE
    KERNEL'S OWN gradient must match the kernel's state to 1e-6 -- nu is quadratic in the gradient, so against the
    oracle's gradient it can only be held to twice the gradient bar (NU_RTOL in the golden tests)."""
    eng, ora, rng = make_pair(D=8, A=4, B=64, kind="adamw", lr=1e-3, seed=41)
    for step in range(3):
        p0 = eng.get_params(0, 0)
        cnt, mu0, nu0 = eng.get_opt_state(0)
        got = eng.train_step_debug()
        want_p, want_s = O.adam_update(p0, got["grads"], {"count": cnt, "mu": mu0, "nu": nu0}, O.OptSpec("adamw", 1e-3))
        p1 = eng.get_params(0, 0)
        cnt1, mu1, nu1 = eng.get_opt_state(0)
        assert int(cnt1) == int(want_s["count"]) == step + 1
        for m in O.MODULES:
            for k in ("w", "b"):
                assert_close(mu1[m][k], want_s["mu"][m][k], rtol=1e-6, atol_scale=1e-7, what=f"mu {m}/{k}")
                assert_close(nu1[m][k], want_s["nu"][m][k], rtol=1e-6, atol_scale=1e-7, what=f"nu {m}/{k}")
                assert_close(p1[m][k], want_p[m][k], rtol=1e-6, atol_scale=1e-7, what=f"theta {m}/{k}")


def test_denormal_second_moment():
    """Adam with a DENORMAL nu (the quotient uses sqrt.approx.ftz / rcp.approx.ftz, which flush denormal inputs): a dead
    hidden unit has an exactly-zero gradient, its nu decays through the denormal range.  sqrt(denormal) < 1.1e-19 is far
    below eps = 1e-8, so flushing it changes the update by < 1e-11 relative: the kernel must agree with the oracle."""
    eng, ora, rng = make_pair(D=8, A=4, B=64, kind="adam", lr=1e-3, seed=43)
    p = eng.get_params(0, 0)
    p[O.MODULES[0]]["b"][0] = -50.0                      # unit 0 of layer 1 never fires: dW2[0, :] == 0 exactly
    eng.set_params(p, 0, 0); ora.params = O.tree_copy(p)
    cnt, mu, nu = eng.get_opt_state(0)
    mu[O.MODULES[1]]["w"][0, :] = np.float32(1e-12)
    nu[O.MODULES[1]]["w"][0, :32] = np.float32(1e-42)    # denormal
    nu[O.MODULES[1]]["w"][0, 32:] = np.float32(2e-38)    # normal now, denormal after a few decays... of 0.999 (stays normal): both paths
    eng.set_opt_state(0, mu, nu, 0)
    ora.opt_state = {"count": np.int32(0), "mu": O.tree_copy(mu), "nu": O.tree_copy(nu)}
    for step in range(3):
        ref = ora.step()
        got = eng.train_step_debug()
        assert not got["grads"][O.MODULES[1]]["w"][0].any() and not ref["grads"][O.MODULES[1]]["w"][0].any()
        p1 = eng.get_params(0, 0)
        _, mu1, nu1 = eng.get_opt_state(0)
        w_got, w_ref = p1[O.MODULES[1]]["w"][0], ora.params[O.MODULES[1]]["w"][0]
        np.testing.assert_allclose(w_got, w_ref, rtol=1e-6, atol=0)                       # update = lr * m_hat / (~0 + eps)
        assert np.all(np.abs(w_got - p[O.MODULES[1]]["w"][0]) > 1e-8)                     # and it is a real, finite move
        nu_got, nu_ref = nu1[O.MODULES[1]]["w"][0], ora.opt_state["nu"][O.MODULES[1]]["w"][0]
        ulp = np.abs(nu_got.view(np.int32).astype(np.int64) - nu_ref.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1 and (nu_got[:32] > 0).all() and (nu_got[:32] < 1.2e-38).all()   # still denormal, not flushed
        p = p1


@pytest.mark.parametrize("loss", ["l2", "mse"])
def test_l2_loss_extension(loss):
    """loss="l2" (0.5 e^2; the reference only has Huber): self-specified, oracle = dqn_oracle.l2_loss."""
    eng, ora, rng = make_pair(D=8, A=4, B=70, kind="adam", lr=1e-3, seed=47)
    eng.set_loss(loss)
    ora.loss = "l2"
    for step in range(3):
        compare_step(eng, ora, what=f"l2 step{step}")
    eng.set_loss("huber")
    ora.loss = "huber"
    compare_step(eng, ora, what="back to huber")


def test_polyak_target_extension():
    """dqn_polyak_target (the reference only hard-copies): bit-exact against the oracle's
    tau * theta + (1 - tau) * theta^- in fp32; tau = 1 equals the hard sync."""
    eng, ora, rng = make_pair(D=9, A=4, B=64, seed=53)
    for step in range(4):
        compare_step(eng, ora, what=f"polyak step{step}")
        eng.polyak_target(0.005)
        ora.update_target_model(tau=0.005)
        t = eng.get_params(0, 1)
        for m in O.MODULES:
            for k in ("w", "b"):
                assert np.array_equal(t[m][k], ora.target_params[m][k]), f"{m}/{k}"
    eng.polyak_target(1.0)
    t, o = eng.get_params(0, 1), eng.get_params(0, 0)
    for m in O.MODULES:
        assert np.array_equal(t[m]["w"], o[m]["w"]) and np.array_equal(t[m]["b"], o[m]["b"])
    with pytest.raises(dqn_b200.DqnError):
        eng.polyak_target(1.5)


def test_agent_tau_and_loss_kwargs(golden):
    """Agent(..., tau=, loss=): _update_target_model becomes a Polyak step; session mode serves the rest."""
    import asyncio
    theta0 = golden_tree(golden["ref_checkpoint"], "params")
    opt = dqn_b200.adamw(2e-4)
    agent = dqn_b200.Agent(network=dqn_b200.Model(4), params=theta0, optimizer=opt, opt_state=opt.init(theta0), env=None,
                           buffer_size=500, obs_shape=(500, 9), ac_shape=(500,), gamma=0.99, epsilon=1.0, epsilon_decay_rate=0.99,
                           min_epsilon=0.15, max_episodes=1, max_steps=10, training_start=64, batch_size=64, train_frequency=4,
                           back_up_frequency=50, replace_frequency=20, reward_to_reach=230.0, num_actions=4,
                           saving_directory="/tmp/dqn_b200_tau", seed=3, tau=0.01, loss="l2")
    ora = OracleAgent(theta0, O.init_opt_state(theta0), O.OptSpec("adamw", 2e-4), 500, 9, 0.99, 64, seed=3, loss="l2")
    data = synthetic_transitions(np.random.default_rng(1), 300, 9, 4, done_p=0.2)
    agent._replay_buffer.add_many(*data)
    ora.replay.add_many(*data)
    for _ in range(3):
        agent._step()
        ora.step()
        asyncio.run(agent._update_target_model())
        ora.update_target_model(tau=0.01)
    t = agent._target_params
    for m in O.MODULES:
        assert_close(t[m]["w"], ora.target_params[m]["w"], what=f"target {m}")
    assert not np.array_equal(t[O.MODULES[1]]["w"], agent._params[O.MODULES[1]]["w"])


def test_add_accepts_strided_observation():
    """ReplayBuffer.add with a non-contiguous observation: the optional C staging helper refuses the buffer, the numpy path
    takes over -- identical ring contents either way."""
    rb = dqn_b200.ReplayBuffer(16, (16, 9), (16,))
    base = np.arange(36, dtype=np.float32)
    rb.add(base[::4], 2, 0.5, base[1::4], True)
    rb.add(base[:9].astype(np.float64), np.int64(1), np.float32(-1.0), base[9:18], False)
    s, a, r, s2, d = (np.asarray(x) for x in (rb.states, rb.actions, rb.rewards, rb.observations, rb.dones))
    assert np.array_equal(s[0], base[::4]) and np.array_equal(s2[0], base[1::4]) and a[0] == 2 and r[0] == 0.5 and d[0]
    assert np.array_equal(s[1], base[:9]) and a[1] == 1 and r[1] == -1.0 and not d[1] and rb.size == 2
