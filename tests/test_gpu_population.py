"""Population mode: many independent agents in one launch, each equal to its own oracle run, and a
sharded population (two handles with agent_id_base, as two ranks would hold) bit-identical to the
unsharded one -- there is no data-path communication to perturb anything."""
import numpy as np
import pytest

import dqn_b200
from conftest import assert_close
from oracle import dqn_oracle as O
from oracle.agent_oracle import OracleAgent
from oracle.replay_oracle import synthetic_transitions

pytestmark = pytest.mark.gpu


def fill(pop, n_fill, D):
    data = []
    for i in range(pop.n_local):
        d = synthetic_transitions(np.random.default_rng([77, pop.global_id(i)]), n_fill, D, 4, done_p=0.1)
        pop.store(i, *d)
        data.append(d)
    return data


def test_population_agents_match_their_oracles():
    n, D, N = 12, 8, 600
    pop = dqn_b200.Population(n, D, 4, N, dqn_b200.adam(1e-3), seed=13)
    data = fill(pop, 500, D)
    oras = []
    for i in range(n):
        hp = pop.hparams[i]
        theta = dqn_b200.unflatten_tree(pop.params_flat(i), D, 4)
        o = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-3), N, D, hp["gamma"],
                        hp["batch_size"], seed=13, agent_id=i)
        o.replay.add_many(*data[i])
        oras.append(o)
    assert len({hp["batch_size"] for hp in pop.hparams}) > 3          # ragged batch sizes in one launch
    pop.train_steps(1)
    pop.train_steps(2)
    for i, o in enumerate(oras):
        for _ in range(3):
            o.step()
        got = dqn_b200.unflatten_tree(pop.params_flat(i), D, 4)
        for m in O.MODULES:
            for k in ("w", "b"):
                assert_close(got[m][k], o.params[m][k], what=f"agent {i} {m}/{k}")


def test_sharded_population_is_bitwise_identical():
    n, D, N = 10, 8, 400
    full = dqn_b200.Population(n, D, 4, N, dqn_b200.adam(1e-3), seed=3)
    fill(full, 300, D)
    shards = [dqn_b200.Population(n, D, 4, N, dqn_b200.adam(1e-3), rank=r, world_size=3, seed=3) for r in range(3)]
    for s in shards:
        fill(s, 300, D)
    full.train_steps(5)
    for s in shards:
        s.train_steps(5)
    for s in shards:
        for i in range(s.n_local):
            assert np.array_equal(s.params_flat(i), full.params_flat(s.global_id(i)))
