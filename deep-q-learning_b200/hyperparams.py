"""Drop-in for ``General/QLearning/hyperparameter_optimization.py`` (the sweep surface).

``ParamAgent`` keeps the reference's constructor defaults and ``inject`` signature
(``hyperparameter_optimization.py:14-91``).  Two reference quirks are handled explicitly:

* F12 -- in the reference the discount used by the jitted target function is captured at
  construction (``q_agent.py:111``), so ``inject(gamma=...)`` only rebinds the never-read
  ``self._gamma`` and the sweep trains with the constructor default ``gamma=0.``.
  ``gamma_mode="frozen"`` (default) reproduces that; ``gamma_mode="live"`` forwards the injected
  discount to the device (what the sweep intends).
* the ``max_episodes`` getter recurses forever in the reference (``:68-70``); here it returns the value.

``optimize`` keeps the reference's loop (suggest -> int casts -> inject -> training -> evaluate ->
register, ``:113-136``).  ``bayes_opt`` is a third-party control-plane dependency that is not
installed here; when it is missing the proposals come from a seeded uniform draw over the same
bounds (``SWEEP_BOUNDS``).
"""
import numpy as np

from .agent import Agent

SWEEP_BOUNDS = {                       # hyperparameter_optimization.py:115-123
    "gamma": [0.9, 0.999],
    "epsilon": [0.6, 1.],
    "epsilon_decay_rate": [0.9, 0.999],
    "min_epsilon": [0.001, 0.2],
    "replace_frequency": [20, 70],
    "batch_size": [38, 70],
    "train_frequency": [2, 15],
}
_INT_KEYS = ("replace_frequency", "batch_size", "train_frequency")


def sample_sweep_point(rng):
    """One point of the sweep space, int-cast like ``optimize`` does (``:128-130``)."""
    p = {k: float(rng.uniform(lo, hi)) for k, (lo, hi) in SWEEP_BOUNDS.items()}
    for k in _INT_KEYS:
        p[k] = int(p[k])
    return p


class ParamAgent(Agent):
    def __init__(self, network, params, optimizer, opt_state, env, buffer_size, obs_shape, ac_shape,
                 max_episodes, max_steps, training_start, back_up_frequency, reward_to_reach, num_actions,
                 saving_directory, monitoring=False, gamma=0., epsilon=0., epsilon_decay_rate=0.,
                 min_epsilon=0., replace_frequency=0, batch_size=0, train_frequency=0,
                 *, gamma_mode="frozen", device=0, seed=0, session=False):
        if gamma_mode not in ("frozen", "live"):
            raise ValueError("gamma_mode must be 'frozen' (reference behaviour) or 'live'")
        self._gamma_mode = gamma_mode
        super().__init__(network, params, optimizer, opt_state, env, buffer_size, obs_shape, ac_shape,
                         gamma, epsilon, epsilon_decay_rate, min_epsilon, max_episodes, max_steps,
                         training_start, batch_size, train_frequency, back_up_frequency, replace_frequency,
                         reward_to_reach, num_actions, saving_directory, monitoring, verbose=0,
                         device=device, seed=seed, session=session)

    @property
    def max_episodes(self):
        return self._max_episodes

    @max_episodes.setter
    def max_episodes(self, value):
        self._max_episodes = value

    def inject(self, gamma, epsilon, epsilon_decay_rate, min_epsilon, replace_frequency, batch_size,
               train_frequency):
        self._gamma = gamma
        if self._gamma_mode == "live":
            self._engine.set_hparams(0, gamma=float(gamma))
        self._epsilon = epsilon
        self._epsilon_decay_rate = epsilon_decay_rate
        self._min_epsilon = min_epsilon
        self._replace_frequency = replace_frequency
        self._batch_size = batch_size          # property: also updates the device copy
        self._train_frequency = train_frequency


def generate_util_func(agent, episodes):
    agent.max_episodes = episodes

    def util_func(gamma, epsilon, epsilon_decay_rate, min_epsilon, replace_frequency, batch_size,
                  train_frequency):
        agent.inject(gamma, epsilon, epsilon_decay_rate, min_epsilon, replace_frequency, batch_size,
                     train_frequency)
        agent.training()
        return agent.evaluate()

    return util_func


def optimize(agent, episodes=500, runs=20):
    black_box_func = generate_util_func(agent, episodes)
    try:
        from bayes_opt import BayesianOptimization, UtilityFunction
        optimizer = BayesianOptimization(black_box_func, pbounds=SWEEP_BOUNDS, verbose=2, random_state=1000)
        util = UtilityFunction(kind="ucb", kappa=1.96, xi=0.01)
        suggest = lambda: optimizer.suggest(util)
        register = optimizer.register
        best = lambda: optimizer.max
    except ImportError:
        rng = np.random.default_rng(1000)
        history = []
        suggest = lambda: sample_sweep_point(rng)
        register = lambda p, t: history.append({"target": t, "params": p})
        best = lambda: max(history, key=lambda e: e["target"])
    for run in range(runs):
        next_point = suggest()
        for k in _INT_KEYS:
            next_point[k] = int(next_point[k])
        target = black_box_func(**next_point)
        register(next_point, target)
        print("Run: {}".format(run))
        print("params: {} \n target: {}".format(next_point, target))
        print("-----")
    return best()
