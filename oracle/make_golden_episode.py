"""Generate ``tests/golden/episode_ref_*.npz`` by RUNNING the reference's own ``Agent.training()`` loop (build container).

    python oracle/make_golden_episode.py        # needs /root/reference (read-only) and numba

``General/QLearning/q_agent.py`` is imported unmodified (its jax / haiku / optax / gym / matplotlib imports resolve to
``oracle/ref_shims``) and ``Agent.training()`` runs against a scripted environment whose rewards and dones come from a
seeded stream and do not depend on the actions, so the loop's CONTROL FLOW -- which the device-side episode kernels
restate -- is reproducible even though the reference's action and sampler RNGs are not seedable per agent:

  * at which env steps ``_step()`` ran                     (q_agent.py:186-187: size >= training_start and step_count % train_frequency == 0)
  * after which episodes ``_update_target_model()`` ran    (q_agent.py:192-193)
  * epsilon after every episode                            (q_agent.py:120-121, :202)
  * episode rewards / the 50-entry window / its average    (q_agent.py:123-126, :203, :219)
  * episode lengths (done, forced done at max_steps :179-180, exhaustion of the step loop bounded by max_episodes :174)
  * where ``training()`` stopped                           (q_agent.py:211, :219-222)

``_step`` and ``_update_target_model`` are observed by wrapping the bound methods on the instance; the wrapped originals
still run (the real numba sampler and the reference's jitted closures over the shims).
"""
import os
import random
import sys
import tempfile

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shims"))

import haiku as hk                                          # noqa: E402  (shim)
import optax                                                # noqa: E402  (shim)
from General.QLearning.q_agent import Agent                 # noqa: E402  (the reference's own code)
from LunarLander.dddqn import Model                         # noqa: E402  (the reference's own code)

D, A = 9, 4


class _Space:
    n = A


class ScriptedEnv:
    """Old-gym API of LunarLander/env.py: reset() -> obs[1, D]; step(a) -> (obs[1, D], reward, done, info).  Rewards are
    multiples of 1/4 (exact in float32 and float64), dones Bernoulli(p): both drawn up front, independent of the action."""
    action_space = _Space()

    def __init__(self, seed, n, done_p):
        rng = np.random.default_rng(seed)
        self.obs = rng.standard_normal((n + 1, 1, D)).astype(np.float32)
        self.rewards = (rng.integers(-8, 13, n) / 4.0).astype(np.float64)
        self.dones = rng.random(n) < done_p
        self.t = 0                                           # env steps taken so far

    def reset(self):
        return self.obs[self.t]

    def step(self, action):
        assert 0 <= int(action) < A
        r, d = float(self.rewards[self.t]), bool(self.dones[self.t])
        self.t += 1
        return self.obs[self.t], r, d, {}


def run_case(name, seed, done_p, **kw):
    random.seed(seed)
    np.random.seed(seed)
    cfg = dict(buffer_size=64, gamma=0.9, epsilon=0.9, epsilon_decay_rate=0.9, min_epsilon=0.05, max_episodes=40, max_steps=9,
               training_start=20, batch_size=8, train_frequency=3, back_up_frequency=1000, replace_frequency=2,
               reward_to_reach=1e9)
    cfg.update(kw)
    n_max = cfg["max_episodes"] * max(cfg["max_steps"], 1) + 8
    env = ScriptedEnv(seed, n_max, done_p)
    model = hk.without_apply_rng(hk.transform(lambda *args: Model(A)(*args)))             # Test/lunar_lander.py:47
    params = model.init(seed, env.reset())
    optimizer = optax.adam(1e-3)
    agent = Agent(network=model, params=params, optimizer=optimizer, opt_state=optimizer.init(params), env=env,
                  obs_shape=(cfg["buffer_size"], D), ac_shape=(cfg["buffer_size"],), num_actions=A,
                  saving_directory=tempfile.mkdtemp(prefix="dqn_ref_golden_"), monitoring=False, verbose=0, **cfg)
    trained_at, synced_after, eps_after, episode_len, step_marks = [], [], [], [], [0]
    real_step, real_sync, real_eps = agent._step, agent._update_target_model, agent._update_epsilon

    def step_spy():
        trained_at.append(env.t)                             # env steps taken when _step() runs (1-based index of the step)
        real_step()

    async def sync_spy():
        synced_after.append(len(eps_after))                  # index of the episode that just ended
        await real_sync()

    async def eps_spy():
        await real_eps()
        eps_after.append(agent._epsilon)
        episode_len.append(env.t - step_marks[-1])
        step_marks.append(env.t)

    agent._step, agent._update_target_model, agent._update_epsilon = step_spy, sync_spy, eps_spy
    agent.training()
    episodes = len(eps_after)
    out = dict(rewards=env.rewards[:env.t], dones=env.dones[:env.t], observations=env.obs[:env.t + 1, 0],
               trained_at=np.array(trained_at, np.int64), synced_after=np.array(synced_after, np.int64),
               eps_after=np.array(eps_after, np.float64), episode_len=np.array(episode_len, np.int64),
               reward_history=np.array(agent._reward_history, np.float64), average_reward=np.float64(agent._average_reward()),
               episodes=np.int64(episodes), env_steps=np.int64(env.t),
               stopped_early=np.bool_(episodes < cfg["max_episodes"]),
               buffer_size_final=np.int64(agent._replay_buffer.size),
               ring_rewards=np.array(agent._replay_buffer.rewards), ring_dones=np.array(agent._replay_buffer.dones),
               **{"cfg_" + k: np.float64(v) for k, v in cfg.items()})
    np.savez_compressed(os.path.join(OUT, f"episode_ref_{name}.npz"), **out)
    print(f"episode_ref_{name}.npz: {episodes} episodes, {env.t} env steps, {len(trained_at)} train steps, "
          f"{len(synced_after)} syncs, stopped early: {bool(out['stopped_early'])}")


def main():
    run_case("basic", 11, 0.12)                                                      # dones, forced dones at max_steps = 9, gate, syncs
    run_case("early_stop", 12, 0.2, reward_to_reach=3.0, max_episodes=60)            # training() returns when the average passes
    run_case("loop_bound", 13, 0.0, max_steps=1500, max_episodes=6, training_start=5, train_frequency=2, replace_frequency=1)   # :174 (sic)
    run_case("window", 14, 0.5, max_episodes=70, max_steps=4, replace_frequency=7)   # more than 50 episodes: the window pops


if __name__ == "__main__":
    main()
