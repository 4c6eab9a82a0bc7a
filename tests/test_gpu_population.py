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


def test_launch_order_changes_nothing():
    """More agents than SMs: the tensor-core population kernel starts the costly agents (batch > 64) first (api.cu,
    train_common).  Agents are independent, so one launch of all 200 must equal -- bit for bit -- the same agents trained in
    four launches of 50 (<= SM count: launched in index order)."""
    n, D, N = 200, 8, 300
    one = dqn_b200.Population(n, D, 4, N, dqn_b200.adam(1e-3), seed=5)
    four = dqn_b200.Population(n, D, 4, N, dqn_b200.adam(1e-3), seed=5)
    fill(one, 200, D)
    fill(four, 200, D)
    sizes = {hp["batch_size"] for hp in one.hparams}
    assert min(sizes) <= 64 < max(sizes)                      # both bodies of the kernel in the one launch
    one.train_steps(3)
    for b in range(0, n, 50):
        four.engine.train_steps(3, agent_begin=b, agent_end=b + 50)
    for i in range(n):
        assert np.array_equal(one.params_flat(i), four.params_flat(i)), f"agent {i} (batch {one.hparams[i]['batch_size']})"


@pytest.mark.timeout(600)
def test_baseline_population_spot_check():
    """BASELINE configs[2] at full size: 1024 sweep agents x 40 000-slot rings (3.9 GB) in ONE launch; 16 agents spread over
    the population (incl. the first, the last, and agents with batch > 64, i.e. two 64-row tiles) are each compared with
    their own oracle run.  Rings of the other agents are filled on the device (their contents are not checked)."""
    import ctypes as C
    import torch
    n, D, N = 1024, 8, 40_000
    pop = dqn_b200.Population(n, D, 4, N, dqn_b200.adam(1e-4), seed=21)
    two_tile = [i for i, hp in enumerate(pop.hparams) if hp["batch_size"] > 64]
    spots = sorted(set([0, 1, 511, 1022, 1023] + two_tile[:4] + two_tile[-3:] + list(range(97, 1024, 181))))[:20]
    assert len(spots) >= 16 and len(set(spots) & set(two_tile)) >= 4
    lib, eng, chk = pop.engine.lib, pop.engine, dqn_b200.pkg._lib.check
    g = torch.Generator(device="cuda")
    dev = lambda t: C.c_void_p(t.data_ptr())
    data = {}
    for i in range(n):
        if i in spots:
            data[i] = synthetic_transitions(np.random.default_rng([78, i]), N, D, 4, done_p=0.05)
            pop.store(i, *data[i])
        else:
            g.manual_seed(5000 + i)
            s = torch.randn(N, D, generator=g, device="cuda"); s2 = torch.randn(N, D, generator=g, device="cuda")
            r = torch.randn(N, generator=g, device="cuda")
            a = torch.randint(0, 4, (N,), generator=g, device="cuda", dtype=torch.int64)
            d = (torch.rand(N, generator=g, device="cuda") < 0.05).to(torch.uint8)
            chk(lib.dqn_store_device(eng.h, i, N, dev(s), dev(a), dev(r), dev(s2), dev(d)))
    oras = {}
    for i in spots:
        hp = pop.hparams[i]
        theta = dqn_b200.unflatten_tree(pop.params_flat(i), D, 4)
        o = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec("adam", 1e-4), N, D, hp["gamma"], hp["batch_size"], seed=21, agent_id=i)
        o.replay.states[:], o.replay.actions[:], o.replay.rewards[:], o.replay.observations[:], o.replay.dones[:] = data[i]
        o.replay.counter = o.replay.size = N
        oras[i] = o
    pop.train_steps(4)                       # one launch: 1024 CTAs x 4 fused steps
    pop.sync_targets()
    pop.train_steps(2)
    for i, o in oras.items():
        for _ in range(4):
            o.step()
        o.update_target_model()
        for _ in range(2):
            o.step()
        got = dqn_b200.unflatten_tree(pop.params_flat(i), D, 4)
        for m in O.MODULES:
            for k in ("w", "b"):
                assert_close(got[m][k], o.params[m][k], what=f"agent {i} (batch {pop.hparams[i]['batch_size']}) {m}/{k}")
        assert abs(float(pop.engine.losses(1, agent=i)[0]) - float(o.last["loss"])) <= 1e-5 * abs(float(o.last["loss"]))
