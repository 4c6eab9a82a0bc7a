"""CPU restatement of the reference's episode loop around the train step (oracle; test infra only).

Follows ``General/QLearning/q_agent.py``: ``_policy`` ``:137-141``, one iteration of ``_run_episode``
``:174-189`` (forced done at ``max_steps`` ``:179-180``, ``replay_buffer.add`` ``:182``, reward accumulation
``:184``, the train gate ``:186-187``, ``break`` on done ``:189`` or exhaustion of the step loop, whose bound is
``max_episodes`` -- sic, ``:174``), the episode epilogue ``:192-203`` (hard sync when ``episode %
replace_frequency == 0``, epsilon decay ``:120-121``, the 50-entry reward window ``:123-126``) and the stop test
of ``training`` ``:211, :219``.

Pinned by ``tests/golden/episode_ref_*.npz``: traces of the reference's own ``Agent.training()`` run over
``oracle/ref_shims`` with a scripted environment (``oracle/make_golden_episode.py``).

Two things cannot follow the reference literally and are stated here instead:
* the reference draws ``random.uniform(0, 1)`` and ``numpy.random.randint(0, A)`` from unseeded host RNGs
  (SURVEY 3.3).  The draws of policy call ``c`` of agent ``g`` are Philox4x32-10 outputs with counter
  ``(0, c_lo, c_hi, g)`` and key ``(seed_lo, seed_hi ^ 0x504F4C49)``: ``u = ((o0 >> 5) * 2**26 + (o1 >> 6)) /
  2**53`` (CPython's ``random()`` construction) and ``randint = (o2 * A) >> 32``;
* ``statistics.mean`` sums exactly (fractions); the window average here is a left-to-right float64 sum divided
  by the length, oldest entry first (what the device does).  The two differ by at most an ulp of the average.
"""
import numpy as np

from .philox import philox4x32_10

POLICY_TAG = 0x504F4C49
WINDOW = 50                     # q_agent.py:125


def policy_draw(seed, agent, call, num_actions):
    """(u, random_action) of policy call ``call`` of ``agent``."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    call = int(call)
    o = philox4x32_10(0, call & 0xFFFFFFFF, (call >> 32) & 0xFFFFFFFF, int(agent) & 0xFFFFFFFF,
                      seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ POLICY_TAG)
    o = [int(np.asarray(x).reshape(-1)[0]) for x in o]
    u = ((o[0] >> 5) * 67108864.0 + (o[1] >> 6)) / 9007199254740992.0
    return u, (o[2] * int(num_actions)) >> 32


class EpisodeOracle:
    """Episode-loop state of one agent; ``agent`` is an ``OracleAgent`` (params, replay ring, ``step()``)."""

    def __init__(self, agent, epsilon, epsilon_decay_rate, min_epsilon, max_episodes, max_steps, training_start,
                 train_frequency, replace_frequency, reward_to_reach, num_actions, seed=0, agent_id=0):
        self.agent = agent
        self.epsilon, self.epsilon_decay_rate, self.min_epsilon = float(epsilon), float(epsilon_decay_rate), float(min_epsilon)
        self.max_episodes, self.max_steps, self.training_start = int(max_episodes), int(max_steps), int(training_start)
        self.train_frequency, self.replace_frequency = int(train_frequency), int(replace_frequency)
        self.reward_to_reach, self.num_actions = float(reward_to_reach), int(num_actions)
        self.seed, self.agent_id = seed, agent_id
        self.policy_calls = 0
        self.step_count = 0            # q_agent.py:210
        self.episode = 0               # q_agent.py:211
        self.step_in_episode = 0       # `step` of q_agent.py:174 (counts from 1)
        self.epi_reward = 0.0          # q_agent.py:172
        self.reward_history = []       # q_agent.py:106
        self.average_reward = 0.0
        self.finished = False

    def policy(self, state):
        """q_agent.py:137-141 -> (action, took_greedy_branch)."""
        u, rnd = policy_draw(self.seed, self.agent_id, self.policy_calls, self.num_actions)
        self.policy_calls += 1
        if self.epsilon < u:
            return int(self.agent.policy_greedy(np.asarray(state, np.float32).reshape(1, -1))), True
        return int(rnd), False

    def observe(self, state, action, reward, observation, done):
        """q_agent.py:175-203 for one env step (after env.step); returns what happened."""
        step = self.step_in_episode + 1
        self.step_count += 1                                                    # :175
        done = bool(done)
        if step == self.max_steps:                                              # :179-180
            done = True
        self.agent.add(state, action, reward, observation, done)                # :182
        self.epi_reward += float(np.float32(reward))                            # :184
        trained = False
        if self.agent.replay.size >= self.training_start and self.step_count % self.train_frequency == 0:   # :186
            self.agent.step()                                                   # :187
            trained = True
        ended = done or step == self.max_episodes                               # :189 / loop bound of :174
        synced = False
        if ended:
            if self.episode % self.replace_frequency == 0:                      # :192-193
                self.agent.update_target_model()
                synced = True
            self.epsilon = max(self.epsilon * self.epsilon_decay_rate, self.min_epsilon)    # :120-121, :202
            self.reward_history.append(self.epi_reward)                         # :123-126, :203
            while len(self.reward_history) > WINDOW:
                self.reward_history.pop(0)
            total = 0.0
            for x in self.reward_history:
                total += x
            self.average_reward = total / len(self.reward_history)
            self.epi_reward = 0.0
            self.step_in_episode = 0
            self.episode += 1
            if self.average_reward > self.reward_to_reach or self.episode >= self.max_episodes:   # :219, :211
                self.finished = True
        else:
            self.step_in_episode = step
        return dict(ended=ended, trained=trained, synced=synced, done=done)
