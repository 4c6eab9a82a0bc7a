"""CPU restatement of the hot-path half of ``Agent`` (oracle; test infra only).

Follows ``General/QLearning/q_agent.py``: constructor state ``:87-113`` (``target_params = params``
alias at ``:91``), ``_policy`` greedy branch ``:137-139``, ``_update_target_model`` ``:143-144``
(hard copy), ``_step`` ``:146-169``, and the ``replay_buffer.add`` call site ``:182``.  The
sampler's RNG is replaced by explicit indices or the Philox convention of ``oracle/philox.py``
(SURVEY F8): everything downstream of the indices is the reference's arithmetic.
"""
import numpy as np

from . import dqn_oracle as O
from .philox import sample_indices
from .replay_oracle import OracleReplay, gather


class OracleAgent:
    def __init__(self, params, opt_state, opt, buffer_size, obs_dim, gamma, batch_size,
                 seed=0, agent_id=0, loss="huber"):
        self.params = O.tree_copy(params)
        self.target_params = O.tree_copy(params)            # q_agent.py:91
        self.opt_state = {"count": np.int32(opt_state["count"]),
                          "mu": O.tree_copy(opt_state["mu"]), "nu": O.tree_copy(opt_state["nu"])}
        self.opt = opt
        self.replay = OracleReplay(buffer_size, (buffer_size, obs_dim), (buffer_size,))
        self.gamma = gamma
        self.batch_size = batch_size
        self.seed, self.agent_id = seed, agent_id
        self.train_steps = 0
        self.last = None
        self.loss = loss                                            # "huber" = the reference; "l2" = extension (dqn_oracle.l2_loss)

    def add(self, state, action, reward, observation, done):        # q_agent.py:182
        self.replay.add(state, action, reward, observation, done)

    def policy_greedy(self, state):                                 # q_agent.py:139
        return O.compute_action(self.params, state)

    def update_target_model(self, tau=None):                        # q_agent.py:143-144 (tau: Polyak extension, dqn_oracle.polyak)
        self.target_params = O.tree_copy(self.params) if tau is None else O.polyak(self.target_params, self.params, tau)

    def step(self, indices=None, device_h=None, tie_tol=0.0):      # q_agent.py:146-169 (device_h: see dqn_oracle.relu_masks)
        if indices is None:
            indices = sample_indices(self.seed, self.agent_id, self.train_steps,
                                     self.batch_size, self.replay.size)
        batch = gather(indices, *self.replay.arrays())
        self.params, self.opt_state, parts = O.train_step(
            self.params, self.target_params, self.opt_state, batch, self.gamma, self.opt,
            return_parts=True, loss=self.loss, device_h=device_h, tie_tol=tie_tol)
        parts["indices"] = np.asarray(indices, dtype=np.int64)
        self.last = parts
        self.train_steps += 1
        return parts
