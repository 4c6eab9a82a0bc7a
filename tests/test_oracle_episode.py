"""oracle/episode_oracle.py: properties of the restated episode loop (q_agent.py:137-141, :171-222) that can be
checked without the reference's runtime (its host RNGs are unseeded)."""
import numpy as np

from oracle import dqn_oracle as O
from oracle.agent_oracle import OracleAgent
from oracle.episode_oracle import EpisodeOracle, policy_draw


def make(**kw):
    theta = O.init_params(np.random.default_rng(0), 4, 3, bias_std=0.05)
    ag = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-3), 40, 4, 0.9, 8, seed=3, agent_id=2)
    cfg = dict(epsilon=0.5, epsilon_decay_rate=0.9, min_epsilon=0.2, max_episodes=100, max_steps=5, training_start=10,
               train_frequency=3, replace_frequency=2, reward_to_reach=1e9, num_actions=3, seed=3, agent_id=2)
    cfg.update(kw)
    return EpisodeOracle(ag, **cfg)


def test_policy_draws_are_uniform_and_reproducible():
    us, acts = zip(*[policy_draw(9, 1, c, 4) for c in range(4000)])
    assert 0.0 <= min(us) and max(us) < 1.0 and abs(np.mean(us) - 0.5) < 0.02
    assert np.all(np.bincount(acts, minlength=4) > 900)
    assert policy_draw(9, 1, 17, 4) == policy_draw(9, 1, 17, 4) != policy_draw(9, 2, 17, 4)


def test_epsilon_branches():
    greedy, rnd = make(epsilon=0.0, min_epsilon=0.0), make(epsilon=1.0, epsilon_decay_rate=1.0)
    s = np.ones(4, np.float32)
    assert all(greedy.policy(s)[1] for _ in range(50))          # epsilon 0 < u almost surely
    assert not any(rnd.policy(s)[1] for _ in range(50))         # u < 1 always


def test_episode_bookkeeping():
    e = make()
    rng = np.random.default_rng(1)
    trained = synced = 0
    for t in range(1, 61):
        ev = e.observe(rng.standard_normal(4), 1, 1.0, rng.standard_normal(4), False)
        assert ev["ended"] == (t % 5 == 0) and ev["done"] == (t % 5 == 0)             # forced done at max_steps
        assert ev["trained"] == (t >= 10 and t % 3 == 0)                              # gate: size >= start, every 3rd step
        trained += ev["trained"]
        synced += ev["synced"]
    assert e.episode == 12 and synced == 6 and e.agent.train_steps == trained == 17
    assert e.reward_history == [5.0] * 12 and e.average_reward == 5.0
    assert abs(e.epsilon - max(0.5 * 0.9 ** 12, 0.2)) < 1e-15 and e.epsilon == 0.2
    short = make(max_steps=1000, max_episodes=4)                                     # loop bound is max_episodes (sic)
    ends = [short.observe(np.zeros(4), 0, 0.0, np.zeros(4), False)["ended"] for _ in range(16)]
    assert ends == [False, False, False, True] * 4 and short.finished
    win = make(max_steps=1)
    for i in range(60):
        win.observe(np.zeros(4), 0, float(i), np.zeros(4), False)
    assert len(win.reward_history) == 50 and win.reward_history[0] == 10.0 and win.average_reward == np.mean(np.arange(10, 60))
