"""The whole drop-in loop on the device path: ``Agent.training()`` (q_agent.py:209-222) driving ``_policy`` ->
``env.step`` -> ``ReplayBuffer.add`` -> ``_step`` -> episode epilogue against a small synthetic environment with the
reference's old-gym API (LunarLander/env.py: ``reset() -> obs[1, D]``, ``step(a) -> (obs[1, D], reward, done, info)``).
The environment is a contextual bandit stretched into episodes, so "learning" is checkable in seconds: reward +1 for
the action that matches the largest of the first four observation entries, -0.25 otherwise."""
import numpy as np
import pytest

import dqn_b200

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


class ArgmaxEnv:
    """obs = 8 features + fraction of the episode elapsed (the ObsWrapper convention, env.py:17-21), shape (1, 9) f32."""

    def __init__(self, seed, horizon=20):
        self.rng, self.horizon, self.t = np.random.default_rng(seed), horizon, 0
        self.obs = None

    def _draw(self):
        f = self.rng.standard_normal(8).astype(np.float32)
        self.obs = np.concatenate([f, [self.t / self.horizon]]).astype(np.float32)[None, :]
        return self.obs

    def reset(self):
        self.t = 0
        return self._draw()

    def step(self, action):
        reward = 1.0 if int(action) == int(np.argmax(self.obs[0, :4])) else -0.25
        self.t += 1
        return self._draw(), reward, self.t >= self.horizon, {}


@pytest.mark.parametrize("session", [False, True])
def test_agent_training_learns_on_a_synthetic_env(session, tmp_path):
    import random
    random.seed(0)                       # Agent._policy draws from the host RNGs, like the reference (q_agent.py:138,141)
    np.random.seed(0)
    model = dqn_b200.Model(4)
    rng = np.random.default_rng(0)
    params = model.init(rng, np.zeros((1, 9), np.float32))
    opt = dqn_b200.adam(2e-3)
    agent = dqn_b200.Agent(network=model, params=params, optimizer=opt, opt_state=opt.init(params), env=ArgmaxEnv(1),
                           buffer_size=5000, obs_shape=(5000, 9), ac_shape=(5000,), gamma=0.0, epsilon=1.0,
                           epsilon_decay_rate=0.97, min_epsilon=0.02, max_episodes=150, max_steps=1500, training_start=64,
                           batch_size=64, train_frequency=1, back_up_frequency=1000, replace_frequency=5,
                           reward_to_reach=1e9, num_actions=4, saving_directory=str(tmp_path), verbose=0, seed=3,
                           session=session)
    history = []
    step_count = 0
    for episode in range(150):                                  # Agent.training()'s loop, keeping every episode's reward
        step_count = agent._run_episode(step_count, episode)
        history.append(agent._reward_history[-1])
    first, last = np.mean(history[:10]), np.mean(history[-20:])
    assert first < 7.0                                          # random policy: 20 * (0.25 * 1 - 0.75 * 0.25) = 1.25 on average
    assert last > 14.0, f"no learning: first 10 episodes {first:.2f}, last 20 episodes {last:.2f}"    # optimum = 20
    assert agent._replay_buffer.size == 3000 and agent._engine.train_step_count() == 3000 - 63
    # greedy evaluation straight from the device weights
    env, hits = ArgmaxEnv(99), 0
    s = env.reset()
    for _ in range(200):
        a = agent._compute_action(None, s)
        hits += int(a == int(np.argmax(s[0, :4])))
        s, _, done, _ = env.step(a)
        if done:
            s = env.reset()
    assert hits >= 160
