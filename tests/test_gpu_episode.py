"""Episode-loop control on the device (csrc/episode.cu) against oracle/episode_oracle.py: a
population of agents with different schedules is driven for a few hundred env steps from ONE externally generated
transition stream; the device loop (policy -> observe -> gated train -> hard sync) must take the same decisions as
the restated reference loop.  Bit-exact: epsilon-greedy branch and random actions, stored ring contents, epsilon,
episode / step counters, reward window average, Adam step counts, sync events.  Greedy actions agree wherever the
oracle's Q margin is clear; parameters stay within the drift tolerance of the multi-step train tests."""
import numpy as np
import pytest

import dqn_b200
from conftest import assert_close
from oracle import dqn_oracle as O
from oracle.agent_oracle import OracleAgent
from oracle.episode_oracle import EpisodeOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["cta", "cluster", "cta_tc"])
def step_kernel(request, monkeypatch):
    monkeypatch.setenv("DQN_B200_STEP_KERNEL", request.param)
    return request.param


def test_device_episode_loop_matches_restated_reference_loop():
    import torch
    D, A, N, n_agents, seed = 9, 4, 64, 6, 31
    rng = np.random.default_rng(2)
    eng = dqn_b200.DqnEngine(D, A, N, 16, 0.99, dqn_b200.adam(1e-3), n_agents=n_agents, seed=seed, agent_id_base=10)
    cfgs, oracles = [], []
    for g in range(n_agents):
        theta = O.init_params(np.random.default_rng([7, g]), D, A, bias_std=0.05)
        B = int(rng.integers(8, 24))
        cfg = dict(epsilon=float(rng.uniform(0.3, 1.0)), epsilon_decay_rate=float(rng.uniform(0.8, 0.99)),
                   min_epsilon=float(rng.uniform(0.05, 0.3)), reward_to_reach=1e9, max_episodes=1000,
                   max_steps=int(rng.integers(5, 14)), training_start=int(rng.integers(10, 30)),
                   train_frequency=int(rng.integers(1, 5)), replace_frequency=int(rng.integers(1, 4)))
        if g == 0:
            cfg.update(epsilon=0.0, min_epsilon=0.0)                  # always greedy
        if g == 1:
            cfg.update(epsilon=1.5, epsilon_decay_rate=1.0)           # always random (epsilon never below a draw)
        if g == 2:
            cfg.update(max_steps=1000, max_episodes=9)                # the step loop runs out first (bound = max_episodes, sic)
        eng.set_params(theta, g, 0)
        eng.set_params(theta, g, 1)
        eng.set_hparams(g, gamma=0.9 + 0.01 * g, batch_size=B)
        ag = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-3), N, D, 0.9 + 0.01 * g, B, seed=seed, agent_id=10 + g)
        oracles.append(EpisodeOracle(ag, num_actions=A, seed=seed, agent_id=10 + g, **cfg))
        cfgs.append(cfg)
    eng.configure_episodes(cfgs)

    dev = "cuda:0"
    state = rng.standard_normal((n_agents, D)).astype(np.float32)
    unclear = 0
    for t in range(260):
        acts = eng.policy(torch.from_numpy(state).to(dev)).cpu().numpy()
        obs = rng.standard_normal((n_agents, D)).astype(np.float32)
        rew = (2.0 * rng.standard_normal(n_agents)).astype(np.float32)
        done = rng.random(n_agents) < 0.12
        for g, ora in enumerate(oracles):
            want, greedy = ora.policy(state[g])
            if greedy:
                q = O.forward(ora.agent.params, state[g:g + 1])[0]
                assert q[acts[g]] >= q.max() - 1e-4 * max(1.0, np.abs(q).max()), f"t{t} agent{g}: not a greedy action"
                unclear += int(acts[g] != want)
            else:
                assert acts[g] == want, f"t{t} agent{g}: random action"
        ended = eng.observe(torch.from_numpy(state).to(dev), torch.from_numpy(acts).to(dev), torch.from_numpy(rew).to(dev),
                            torch.from_numpy(obs).to(dev), torch.from_numpy(done.astype(np.uint8)).to(dev))
        eng.train_flagged()
        ended = ended.cpu().numpy().astype(bool)
        for g, ora in enumerate(oracles):
            ev = ora.observe(state[g], int(acts[g]), float(rew[g]), obs[g], bool(done[g]))   # device's action on both sides
            assert ev["ended"] == ended[g], f"t{t} agent{g}: episode end"
        state = np.where(ended[:, None], rng.standard_normal((n_agents, D)).astype(np.float32), obs)   # env.reset()
    assert unclear <= 2                                            # near-ties only
    trained_total = 0
    for g, ora in enumerate(oracles):
        st = eng.episode_state(g)
        assert st["epsilon"] == ora.epsilon, f"agent{g} epsilon"                          # bit-exact doubles
        assert (st["episode"], st["step_in_episode"], st["step_count"], st["policy_calls"]) == \
               (ora.episode, ora.step_in_episode, ora.step_count, ora.policy_calls), f"agent{g} counters"
        assert st["window_len"] == len(ora.reward_history)
        assert st["average_reward"] == ora.average_reward and st["episode_reward"] == ora.epi_reward, f"agent{g} rewards"
        assert st["finished"] == int(ora.finished)
        assert eng.buffer_state(g) == (ora.agent.replay.size, ora.agent.replay.counter)
        rp = ora.agent.replay
        for x, y in zip(eng.buffer_export(g), (rp.states, rp.actions, rp.rewards, rp.observations, rp.dones)):
            assert np.array_equal(x, y), f"agent{g} ring"
        cnt, _, _ = eng.get_opt_state(g)
        assert int(cnt) == int(ora.agent.opt_state["count"]) == ora.agent.train_steps, f"agent{g} train steps"
        trained_total += ora.agent.train_steps
        p, tp = eng.get_params(g, 0), eng.get_params(g, 1)
        for m in O.MODULES:
            for k in ("w", "b"):
                assert_close(p[m][k], ora.agent.params[m][k], rtol=2e-4, what=f"agent{g} param {m}/{k}")
                assert_close(tp[m][k], ora.agent.target_params[m][k], rtol=2e-4, what=f"agent{g} target {m}/{k}")
    assert trained_total > 200 and oracles[2].episode >= 20 and any(o.episode > 25 for o in oracles)
    # a second training() run: counters restart, the reward window and the policy stream continue
    calls = eng.episode_state(3)["policy_calls"]
    eng.configure_episodes(cfgs, reset_counters=True)
    st = eng.episode_state(3)
    assert (st["episode"], st["step_count"], st["window_len"], st["policy_calls"]) == (0, 0, len(oracles[3].reward_history), calls)


def test_finished_flag_and_unconfigured_agents():
    import torch
    eng = dqn_b200.DqnEngine(4, 2, 32, 4, 0.9, dqn_b200.adam(1e-3), n_agents=2, seed=1)
    z = torch.zeros(2, 4, device="cuda:0")
    with pytest.raises(dqn_b200.DqnError):
        eng.policy(z)                                                # not configured
    cfg = dict(epsilon=1.0, epsilon_decay_rate=0.5, min_epsilon=0.1, reward_to_reach=2.5, max_episodes=100, max_steps=3,
               training_start=1000, train_frequency=1, replace_frequency=1)
    eng.configure_episodes([cfg, dict(cfg, max_episodes=2)])
    a = torch.zeros(2, dtype=torch.int32, device="cuda:0")
    r = torch.ones(2, device="cuda:0")
    d = torch.zeros(2, dtype=torch.uint8, device="cuda:0")
    ends = []
    for t in range(6):
        ends.append(eng.observe(z, a, r, z, d).cpu().numpy().tolist())
        eng.train_flagged()
    # agent 0: forced done every max_steps = 3 steps; agent 1: the step loop is bounded by max_episodes = 2 (sic, q_agent.py:174)
    assert ends == [[0, 0], [0, 1], [1, 0], [0, 1], [0, 0], [1, 1]]
    s0, s1 = eng.episode_state(0), eng.episode_state(1)
    assert s0["average_reward"] == 3.0 and s0["finished"] == 1                # average 3.0 > reward_to_reach 2.5
    assert s1["episode"] == 3 and s1["finished"] == 1 and s1["average_reward"] == 2.0   # ran out of episodes (flag is sticky)
    assert s0["epsilon"] == 0.25 and eng.buffer_state(0) == (6, 6)


@pytest.mark.parametrize("case", ["basic", "early_stop", "loop_bound", "window"])
def test_device_loop_follows_the_reference_training_traces(case):
    """The device loop against traces of the reference's own ``Agent.training()`` (tests/golden/episode_ref_*.npz, produced
    by oracle/make_golden_episode.py from the unmodified ``q_agent.py``): same reward / done stream -> it trains at the
    same env steps, ends episodes at the same steps, decays epsilon to the same doubles, keeps the same reward window,
    stores the same ring and raises `finished` where ``training()`` returned."""
    import os
    import torch
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, f"episode_ref_{case}.npz"), allow_pickle=False)
    cfg = {k[4:]: g[k].item() for k in g.files if k.startswith("cfg_")}
    N, B = int(cfg["buffer_size"]), int(cfg["batch_size"])
    eng = dqn_b200.DqnEngine(9, 4, N, B, cfg["gamma"], dqn_b200.adam(1e-3), seed=5)
    theta = O.init_params(np.random.default_rng(0), 9, 4, bias_std=0.05)
    eng.set_params(theta, 0, 0)
    eng.set_params(theta, 0, 1)
    eng.configure_episodes([dict(epsilon=cfg["epsilon"], epsilon_decay_rate=cfg["epsilon_decay_rate"], min_epsilon=cfg["min_epsilon"],
                                 reward_to_reach=cfg["reward_to_reach"], max_episodes=int(cfg["max_episodes"]), max_steps=int(cfg["max_steps"]),
                                 training_start=int(cfg["training_start"]), train_frequency=int(cfg["train_frequency"]),
                                 replace_frequency=int(cfg["replace_frequency"]))])
    dev = "cuda:0"
    obs = torch.from_numpy(np.ascontiguousarray(g["observations"])).to(dev)
    rew = torch.from_numpy(g["rewards"].astype(np.float32)).to(dev)          # multiples of 1/4: exact in float32
    don = torch.from_numpy(g["dones"].astype(np.uint8)).to(dev)
    trained_at, episode_len, eps_after, last_end, trained = [], [], [], 0, 0
    for t in range(int(g["env_steps"])):
        assert eng.episode_state(0)["finished"] == 0, "the reference was still running here"
        a = eng.policy(obs[t:t + 1])
        ended = eng.observe(obs[t:t + 1], a, rew[t:t + 1], obs[t + 1:t + 2], don[t:t + 1])
        eng.train_flagged()
        n = eng.train_step_count(0)
        if n != trained:
            trained_at.append(t + 1)
            trained = n
        if int(ended.cpu()[0]):
            episode_len.append(t + 1 - last_end)
            last_end = t + 1
            eps_after.append(eng.episode_state(0)["epsilon"])
    st = eng.episode_state(0)
    assert trained_at == g["trained_at"].tolist()
    assert episode_len == g["episode_len"].tolist()
    assert np.array_equal(np.array(eps_after), g["eps_after"])                # bit-exact doubles
    assert st["episode"] == int(g["episodes"]) and st["step_count"] == int(g["env_steps"]) and st["finished"] == 1
    assert st["window_len"] == len(g["reward_history"]) and st["last_episode_reward"] == g["reward_history"][-1]
    assert abs(st["average_reward"] - float(g["average_reward"])) <= 4e-16 * max(1.0, abs(float(g["average_reward"])))
    s, a_, r, s2, d = eng.buffer_export(0)
    assert eng.buffer_state(0)[0] == int(g["buffer_size_final"])
    assert np.array_equal(r, g["ring_rewards"].astype(np.float32)) and np.array_equal(d, g["ring_dones"])
    cnt, _, _ = eng.get_opt_state(0)
    assert int(cnt) == len(g["trained_at"])
