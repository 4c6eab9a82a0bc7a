"""Pins oracle/episode_oracle.py to golden traces produced by RUNNING the reference's own ``Agent.training()``
(oracle/make_golden_episode.py: ``General/QLearning/q_agent.py`` imported unmodified, scripted environment).  The
restated loop must take every control-flow decision the reference took on the same reward / done stream: when it
trains, when it hard-syncs, how epsilon decays, how long episodes are, what the reward window holds, where it stops."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import dqn_oracle as O
from oracle.agent_oracle import OracleAgent
from oracle.episode_oracle import EpisodeOracle

CASES = ["basic", "early_stop", "loop_bound", "window"]


@pytest.mark.parametrize("case", CASES)
def test_episode_oracle_follows_the_reference_loop(case):
    g = np.load(os.path.join(GOLDEN, f"episode_ref_{case}.npz"), allow_pickle=False)
    cfg = {k[4:]: g[k].item() for k in g.files if k.startswith("cfg_")}
    theta = O.init_params(np.random.default_rng(0), 9, 4, bias_std=0.05)
    ag = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-3), int(cfg["buffer_size"]), 9, cfg["gamma"],
                     int(cfg["batch_size"]), seed=1)
    e = EpisodeOracle(ag, cfg["epsilon"], cfg["epsilon_decay_rate"], cfg["min_epsilon"], int(cfg["max_episodes"]), int(cfg["max_steps"]),
                      int(cfg["training_start"]), int(cfg["train_frequency"]), int(cfg["replace_frequency"]), cfg["reward_to_reach"], 4, seed=1)
    obs, rew, don = g["observations"], g["rewards"], g["dones"]
    trained_at, synced_after, eps_after, episode_len, last_end = [], [], [], [], 0
    for t in range(int(g["env_steps"])):
        assert not e.finished, "the reference was still running here"
        action, _ = e.policy(obs[t])
        ev = e.observe(obs[t], action, rew[t], obs[t + 1], bool(don[t]))
        if ev["trained"]:
            trained_at.append(t + 1)
        if ev["ended"]:
            if ev["synced"]:
                synced_after.append(e.episode - 1)
            eps_after.append(e.epsilon)
            episode_len.append(t + 1 - last_end)
            last_end = t + 1
    assert trained_at == g["trained_at"].tolist()
    assert synced_after == g["synced_after"].tolist()
    assert np.array_equal(np.array(eps_after), g["eps_after"])                       # bit-exact doubles
    assert episode_len == g["episode_len"].tolist()
    assert e.episode == int(g["episodes"]) and e.step_count == int(g["env_steps"])
    assert np.array_equal(np.array(e.reward_history), g["reward_history"])
    assert abs(e.average_reward - float(g["average_reward"])) <= 4e-16 * max(1.0, abs(float(g["average_reward"])))   # statistics.mean is exact
    assert e.finished                                            # the reference's training() returned after this episode
    assert bool(g["stopped_early"]) == (e.episode < int(cfg["max_episodes"]))
    assert ag.replay.size == int(g["buffer_size_final"])
    assert np.array_equal(ag.replay.rewards, g["ring_rewards"]) and np.array_equal(ag.replay.dones, g["ring_dones"])
    assert ag.train_steps == len(g["trained_at"])


def test_golden_traces_cover_the_branches():
    b, s, l, w = (np.load(os.path.join(GOLDEN, f"episode_ref_{c}.npz")) for c in CASES)
    assert (b["episode_len"] == 9).any() and (b["episode_len"] < 9).any()            # forced done at max_steps and real dones
    assert bool(s["stopped_early"]) and float(s["average_reward"]) > 3.0
    assert (l["episode_len"] == 6).all() and not l["dones"].any()                    # step loop bounded by max_episodes = 6 (sic)
    assert int(w["episodes"]) == 70 and len(w["reward_history"]) == 50
