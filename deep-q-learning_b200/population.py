"""A population of independent agents from the hyper-parameter sweep, one CTA per agent.

The reference runs its sweep strictly sequentially (``hyperparameter_optimization.py:126-135``: one
``ParamAgent``, 20 runs).  Every run is an independent unit -- own parameters, target network, Adam
state, replay ring and hyper-parameters (gamma, batch_size from the bounds at ``:115-123``) -- so a
population of them is batched into a single launch (grid = agents) and sharded over GPUs with no
data-path communication: rank ``r`` of ``W`` owns the contiguous agent block ``shard_range(n, r, W)``.
"""
import numpy as np

from .engine import DqnEngine
from .hyperparams import sample_sweep_point
from .specs import Model, flatten_tree


def shard_range(n_agents, rank, world_size):
    """Contiguous block of global agent ids owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n_agents), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def sweep_hparams(n_agents, seed=1000):
    """Per-agent (gamma, batch_size, ...) drawn from the sweep bounds; agent ``g`` gets the same
    draw on every rank / world size (the stream is indexed by the global agent id)."""
    out = []
    for g in range(n_agents):
        out.append(sample_sweep_point(np.random.default_rng([seed, g])))
    return out


class Population:
    def __init__(self, n_agents_global, obs_dim, num_actions, buffer_size, optimizer, rank=0, world_size=1,
                 seed=0, hparam_seed=1000, device=0, max_batch=70, init_seed=0):
        self.n_global = int(n_agents_global)
        self.begin, self.end = shard_range(self.n_global, rank, world_size)
        self.n_local = self.end - self.begin
        self.obs_dim, self.num_actions = obs_dim, num_actions
        self.hparams = sweep_hparams(self.n_global, hparam_seed)[self.begin:self.end]
        self.engine = DqnEngine(obs_dim, num_actions, buffer_size, max_batch, 0.0, optimizer,
                                n_agents=self.n_local, seed=seed, device=device, agent_id_base=self.begin)
        model = Model(num_actions)
        for i, hp in enumerate(self.hparams):
            g = self.begin + i
            tree = model.init(np.random.default_rng([init_seed, g]), np.zeros((1, obs_dim), np.float32))
            flat = flatten_tree(tree, obs_dim, num_actions)
            self.engine.set_params_flat(flat, i, 0)
            self.engine.set_params_flat(flat, i, 1)
            self.engine.set_hparams(i, gamma=hp["gamma"], batch_size=hp["batch_size"])

    def global_id(self, local):
        return self.begin + local

    def store(self, local, s, a, r, s2, done):
        self.engine.store(s, a, r, s2, done, agent=local)

    def train_steps(self, K=1):
        """K train steps of every local agent in ONE launch (grid = local agents)."""
        self.engine.train_steps(K, agent_begin=0, agent_end=self.n_local)

    def sync_targets(self):
        self.engine.sync_target(0, self.n_local)

    # -- the reference's episode loop, batched on the device (q_agent.py:171-222) -----------------
    def configure_episodes(self, max_episodes, max_steps, training_start, reward_to_reach, reset_counters=True):
        """Per-agent epsilon schedule / cadences from the sweep draw (hyperparameter_optimization.py:115-123), the rest
        from the sweep script's constants (Test/lunar_lander_hyper_params.py:22-30)."""
        cfgs = [dict(epsilon=hp["epsilon"], epsilon_decay_rate=hp["epsilon_decay_rate"], min_epsilon=hp["min_epsilon"],
                     reward_to_reach=reward_to_reach, max_episodes=max_episodes, max_steps=max_steps,
                     training_start=training_start, train_frequency=hp["train_frequency"],
                     replace_frequency=hp["replace_frequency"]) for hp in self.hparams]
        self.engine.configure_episodes(cfgs, 0, reset_counters)

    def policy(self, states, actions_out=None):
        return self.engine.policy(states, actions_out, 0, self.n_local)

    def observe(self, states, actions, rewards, observations, dones, episode_end_out=None):
        return self.engine.observe(states, actions, rewards, observations, dones, episode_end_out, 0, self.n_local)

    def train_flagged(self):
        self.engine.train_flagged(0, self.n_local)

    def params_flat(self, local):
        return self.engine.get_params_flat(local, 0)
