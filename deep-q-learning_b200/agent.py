"""Drop-in for ``General/QLearning/q_agent.py::Agent`` over the fused sm_100a train step.

Constructor keyword names, the hyper-parameter attributes, ``_policy``, ``_step``,
``_update_target_model``, ``_run_episode``, ``training`` and ``evaluate`` keep the reference's
meaning (``q_agent.py:61-231``).  What changes is where the state lives: ``params``,
``target_params``, ``opt_state`` and the replay ring are device-resident behind one ``dqn_handle``;
the ``_params`` / ``_target_params`` / ``_opt_state`` attributes are properties that copy to/from
the device in the reference's tree layout.  The four jitted closures the reference stores on the
agent (``_compute_action``, ``_compute_q_targets``, ``_train_step`` and ``preprocessing``,
``q_agent.py:110-112,154``) are fused: ``_step()`` is one kernel launch.

``session=True`` keeps ONE launch of the train-step kernel resident: ``_policy``'s greedy branch, ``add`` + ``_step`` and
``_update_target_model`` are then served from commands in mapped host memory with no kernel launch on the env loop's
path (``dqn_set_session`` in ``include/dqn_b200.h``); anything else (checkpoints, parameter reads) ends the session first.

``network`` is a ``specs.Model`` and ``optimizer`` a ``specs.Optimizer`` (see ``specs.py``); ``env``
is any object with the old-gym ``reset() -> obs[1,D]`` / ``step(a) -> (obs, reward, done, info)``
API of ``LunarLander/env.py`` and is only needed by the episode loop.
"""
from random import uniform
from statistics import mean

import numpy as np
from numpy.random import randint

from . import _lib
from .checkpoint import generate_saving
from .engine import DqnEngine
from .replay import ReplayBuffer
from .specs import EmptyState, ScaleByAdamState, split_opt_state


class Agent:
    def __init__(self, network, params, optimizer, opt_state, env, buffer_size, obs_shape, ac_shape,
                 gamma, epsilon, epsilon_decay_rate, min_epsilon, max_episodes, max_steps,
                 training_start, batch_size, train_frequency, back_up_frequency, replace_frequency,
                 reward_to_reach, num_actions, saving_directory, monitoring=False, verbose=1,
                 *, device=0, seed=0, session=False, loss="huber", tau=None):
        self._network = network
        self._optimizer = optimizer
        self._env = env
        obs_dim = int(tuple(obs_shape)[1])
        if network.num_actions != num_actions:
            raise ValueError("network.num_actions != num_actions")
        # batch_size 0 is what ParamAgent's constructor passes before inject(); the device needs >= 1
        self._engine = DqnEngine(obs_dim, num_actions, buffer_size, max(int(batch_size), 1), gamma,
                                 optimizer, n_agents=1, seed=seed, device=device, session=session)
        self._lib, self._h = self._engine.lib, self._engine.h
        # Extensions the reference does not have, off by default: loss="l2" trains on 0.5 e^2 instead of
        # Huber(1); tau in (0,1) makes _update_target_model a Polyak step theta^- := tau theta + (1 - tau) theta^-.
        self._tau = None if tau is None else float(tau)
        if loss != "huber":
            self._engine.set_loss(loss)
        self._params = params
        self._opt_state = opt_state
        self._target_params = params                                  # q_agent.py:91
        self._replay_buffer = ReplayBuffer(buffer_size=buffer_size, obs_shape=obs_shape, ac_shape=ac_shape,
                                           _engine=self._engine, _agent=0)
        self._gamma = gamma
        self._epsilon = epsilon
        self._epsilon_decay_rate = epsilon_decay_rate
        self._min_epsilon = min_epsilon
        self._max_episodes = max_episodes
        self._max_steps = max_steps
        self._training_start = training_start
        self._batch_size = batch_size
        self._train_frequency = train_frequency
        self._back_up_frequency = back_up_frequency
        self._replace_frequency = replace_frequency
        self._reward_to_reach = reward_to_reach
        self._num_actions = num_actions
        self._reward_history = []
        self._save_state = generate_saving(saving_directory)
        self._monitoring = monitoring
        self._verbose = verbose
        if self._monitoring:
            self._loss_history = []
            self._episode_losses = []

    # -- device-resident state exposed with the reference's attribute names ----------------------
    @property
    def _params(self):
        return self._engine.get_params(0, _lib.DQN_PARAMS_ONLINE)

    @_params.setter
    def _params(self, tree):
        self._engine.set_params(tree, 0, _lib.DQN_PARAMS_ONLINE)

    @property
    def _target_params(self):
        return self._engine.get_params(0, _lib.DQN_PARAMS_TARGET)

    @_target_params.setter
    def _target_params(self, tree):
        self._engine.set_params(tree, 0, _lib.DQN_PARAMS_TARGET)

    @property
    def _opt_state(self):
        count, mu, nu = self._engine.get_opt_state(0)
        tail = (EmptyState(), EmptyState()) if self._optimizer.kind == "adamw" else (EmptyState(),)
        return (ScaleByAdamState(count, mu, nu),) + tail

    @_opt_state.setter
    def _opt_state(self, opt_state):
        count, mu, nu = split_opt_state(opt_state)
        self._engine.set_opt_state(count, mu, nu, 0)

    # batch_size is read live by the reference's _step (q_agent.py:153): keep the device copy in step
    @property
    def _batch_size(self):
        return self.__batch_size

    @_batch_size.setter
    def _batch_size(self, value):
        self.__batch_size = value
        if int(value) >= 1:
            self._engine.set_hparams(0, batch_size=int(value))

    # -- reference methods ----------------------------------------------------------------------------
    async def _update_epsilon(self):
        self._epsilon = max(self._epsilon * self._epsilon_decay_rate, self._min_epsilon)

    async def _update_reward_history(self, episode_reward):
        self._reward_history.append(episode_reward)
        while len(self._reward_history) > 50:
            self._reward_history.pop(0)

    async def _update_loss_history(self):
        episode_loss = float(np.average(self._episode_losses))
        self._loss_history.append(episode_loss)
        while len(self._episode_losses) > 50:
            self._episode_losses.pop(0)

    def _average_reward(self):
        return mean(self._reward_history)

    def _compute_action(self, params, state):
        """``compute_action(params, state)`` (``q_learning_functions.py:67-73``).  ``params`` is
        accepted for signature compatibility; the device copy of the online parameters is used
        (they are the same object in every reference call site, ``q_agent.py:139,229``)."""
        return self._engine.act(state, 0)

    def _policy(self, state):
        if self._epsilon < uniform(0, 1):
            return int(self._compute_action(None, state))
        else:
            return randint(0, self._num_actions)

    def _sync_target(self):
        if self._tau is None:
            _lib.check(self._lib.dqn_sync_target(self._h, 0, 1))          # q_agent.py:143-144: hard copy
        else:
            self._engine.polyak_target(self._tau, 0, 1)

    async def _update_target_model(self):
        self._sync_target()

    def _step(self, indices=None):
        """One fused train step (sample -> preprocess -> targets -> grad -> adam).  ``indices``
        (optional, i64[batch_size]) replaces the Philox draw -- the hook parity runs use."""
        rb = self._replay_buffer
        if indices is None:
            # the add()s staged since the last step (train_frequency of them, q_agent.py:182-187) ride in the
            # kernel's parameter buffer: one launch, no H2D copy (falls back to store + train for > 16)
            rb.step_staged()
        else:
            rb.flush()
            self._engine.train_steps(1, indices=indices, agent_begin=0, agent_end=1)

    def _steps(self, k):
        """``k`` consecutive ``_step()`` calls in one persistent launch (no stores in between)."""
        self._replay_buffer.flush()
        _lib.check(self._lib.dqn_train_step(self._h, 0, 1, int(k), None, None))

    def _step_debug(self, indices=None):
        self._replay_buffer.flush()
        return self._engine.train_step_debug(indices=indices, agent=0)

    # -- resume (beyond the reference: its checkpoint holds params / opt_state only, General/Base/utils.py:21-29) -------
    def save_resume_state(self, path):
        """Everything the reference's checkpoint omits as well: target network, replay ring, counters, epsilon, rewards."""
        self._replay_buffer.flush()
        st = self._engine.export_state(0)
        st.update(agent_epsilon=np.float64(self._epsilon), agent_reward_history=np.asarray(self._reward_history, np.float64),
                  agent_rb_counter=np.int64(self._replay_buffer._counter), agent_rb_sample_calls=np.int64(self._replay_buffer._sample_calls))
        np.savez_compressed(path, **st)

    def load_resume_state(self, path):
        st = dict(np.load(path, allow_pickle=False))
        self._replay_buffer.flush()
        self._engine.import_state(st, 0)
        self._epsilon = float(st["agent_epsilon"])
        self._reward_history = [float(x) for x in st["agent_reward_history"]]
        rb = self._replay_buffer
        rb._counter = int(st["agent_rb_counter"])
        rb._num_samples = min(rb._counter, rb._buffer_size)
        rb._sample_calls = int(st["agent_rb_sample_calls"])
        self.__dict__["_Agent__batch_size"] = int(st["hparam_batch_size"])

    def _run_episode(self, step_count, episode):
        epi_reward = 0.
        state = self._env.reset()
        for step in range(1, self._max_episodes + 1):     # sic: the reference bounds this by max_episodes (q_agent.py:174)
            step_count += 1
            action = self._policy(state)
            observation, reward, done, info = self._env.step(action)

            if step == self._max_steps:
                done = True

            self._replay_buffer.add(state[0], action, reward, observation[0], done)
            state = observation
            epi_reward += reward

            if self._replay_buffer.size >= self._training_start and step_count % self._train_frequency == 0:
                self._step()
                if self._monitoring:
                    self._episode_losses.append(float(self._engine.losses(1, 0)[0]))

            if done:
                break

        if episode % self._replace_frequency == 0:
            self._sync_target()

        if episode % self._back_up_frequency == 0:
            self._save_state(self._params, self._opt_state)

        if self._monitoring and self._training_start < step_count:
            self._loss_history.append(sum(self._episode_losses))

        self._epsilon = max(self._epsilon * self._epsilon_decay_rate, self._min_epsilon)
        self._reward_history.append(epi_reward)
        while len(self._reward_history) > 50:
            self._reward_history.pop(0)
        return step_count

    def training(self):
        step_count = 0
        for episode in range(self._max_episodes):
            step_count = self._run_episode(step_count, episode)

            if episode % 50 == 0 and self._verbose:
                print("Episode: {} -- Reward: {} -- Average: {}".format(episode, self._reward_history[-1],
                                                                        self._average_reward()))

            if self._average_reward() > self._reward_to_reach:
                self._save_state(self._params, self._opt_state)
                self._plot()
                return

    def evaluate(self):
        evaluation_runs = 10
        for run in range(evaluation_runs):
            state = self._env.reset()
            for step in range(self._max_steps):
                action = int(self._compute_action(None, state))
                state, reward, done, info = self._env.step(action)
        return self._average_reward()

    def _plot(self):
        try:
            import matplotlib.pyplot as plt
        except ImportError:      # plotting is outside the hot path; matplotlib is optional here
            return
        fig = plt.figure()
        plt.plot(list(range(len(self._reward_history))), self._reward_history)
        plt.xlabel("Episodes")
        plt.ylabel("Reward")
        fig.show()
