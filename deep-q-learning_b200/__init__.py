"""deep-q-learning_b200 -- B200-native dueling double-DQN train step + replay path.

Drop-in for the hot path of hal9000universe/deep-q-learning (``Agent._step``, ``ReplayBuffer.add``,
``Agent._policy`` and the hyper-parameter surface); see DESIGN.md and INTEGRATION.md.  The directory
name is not a Python identifier: import it through the ``dqn_b200`` shim at the repository root,
which registers it as the package ``deep_q_learning_b200``.

Reference module                                   -> module here
``General/Base/replay_buffer.py``                   -> ``replay``      (ReplayBuffer, sample_batch)
``General/QLearning/q_agent.py``                    -> ``agent``       (Agent)
``General/QLearning/hyperparameter_optimization.py``-> ``hyperparams`` (ParamAgent, optimize)
``General/QLearning/q_learning_functions.py``,
``LunarLander/dddqn.py``                            -> ``csrc/train_fused.cu`` + ``csrc/act.cu`` (fused kernels)
``General/Base/utils.py`` (pickle layout only)     -> ``checkpoint``
"""
from . import _lib
from ._lib import DqnError
from .specs import Model, Optimizer, adam, adamw, flatten_tree, unflatten_tree, param_count, ScaleByAdamState, EmptyState
from .replay import ReplayBuffer, sample_batch
from .engine import DqnEngine
from .agent import Agent
from .hyperparams import ParamAgent, optimize, generate_util_func, SWEEP_BOUNDS, sample_sweep_point
from .population import Population, shard_range, sweep_hparams
from .large_batch import LargeBatchTrainer
from .per import PrioritizedSampler
from .checkpoint import generate_saving, generate_loading, load_pickle

__all__ = ["Agent", "ParamAgent", "ReplayBuffer", "sample_batch", "Model", "Optimizer", "adam", "adamw",
           "DqnEngine", "DqnError", "Population", "LargeBatchTrainer", "PrioritizedSampler", "shard_range", "sweep_hparams", "optimize",
           "generate_util_func", "generate_saving", "generate_loading", "load_pickle", "flatten_tree",
           "unflatten_tree", "param_count", "SWEEP_BOUNDS", "sample_sweep_point", "ScaleByAdamState", "EmptyState"]
