"""Generate ``tests/golden/train_ref_*.npz`` by EXECUTING the reference's own train-step sources (build container only).

    python oracle/make_golden_train.py          # needs /root/reference (read-only)

``General/QLearning/q_learning_functions.py`` and ``LunarLander/dddqn.py`` are imported unmodified from
``/root/reference``; the third-party modules they import (jax, haiku, optax, gym -- not installable here) resolve to
``oracle/ref_shims`` (see its README for what is restated).  For every case the script drives exactly the body of
``Agent._step`` (``General/QLearning/q_agent.py:154-169``) with the reference's own closures:

    preprocessing(...) -> compute_q_targets(params, target_params, ...) -> train_step(params, opt_state, states, q_targets)

on a seeded synthetic replay content and EXPLICIT minibatch indices (the reference's sampler RNG cannot be seeded,
SURVEY F8), several steps in a row with a hard target sync in between (``q_agent.py:143-144``), and stores the inputs and
everything the reference computed: q-targets, loss (``compute_loss``), gradients (``jax.grad(compute_loss)``), updated
parameters and optimiser state after every step, and greedy actions (``compute_action``).

Nothing in ``tests -m gpu``, ``smoke()`` or ``bench.py`` reads /root/reference; they read the files written here.
"""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))       # jax / haiku / optax / gym
sys.path.insert(0, ROOT)

import haiku as hk                                          # noqa: E402  (shim)
import jax                                                  # noqa: E402  (shim)
import optax                                                # noqa: E402  (shim)
from General.QLearning.q_learning_functions import (        # noqa: E402  (the reference's own code)
    action_computation, generate_loss_computation, generate_q_target_comp, generate_train_step, preprocessing)
from LunarLander.dddqn import Model                         # noqa: E402  (the reference's own code)

from oracle.replay_oracle import synthetic_transitions      # noqa: E402

MODULES = ["model/~/linear", "model/~/linear_1", "model/~/linear_2", "model/~/linear_3"]


class _Space:
    def __init__(self, n):
        self.n = n


class _Env:                                                 # generate_q_target_comp only reads env.action_space.n
    def __init__(self, n):
        self.action_space = _Space(n)


def flat(tree):
    return np.concatenate([np.ravel(np.asarray(tree[m][k], np.float32)) for m in MODULES for k in ("w", "b")])


def run_case(name, D, A, B, N, gamma, opt_kind, lr, steps, sync_at, theta0=None, seed=0, done_p=0.2):
    rng = np.random.default_rng(seed)
    model = hk.without_apply_rng(hk.transform(lambda *args: Model(A)(*args)))            # Test/lunar_lander.py:47
    if theta0 is None:
        params = model.init(seed + 1, np.zeros((1, D), np.float32))
    else:
        params = {m: {k: np.array(theta0[m][k], np.float32) for k in ("w", "b")} for m in MODULES}
    for m in MODULES:                                       # haiku's zero biases hide bias-gradient bugs
        params[m]["b"] = (0.05 * rng.standard_normal(params[m]["b"].shape)).astype(np.float32)
    target_params = {m: {k: (params[m][k] + 0.02 * rng.standard_normal(params[m][k].shape)).astype(np.float32)
                         for k in ("w", "b")} for m in MODULES}
    optimizer = optax.adamw(lr) if opt_kind == "adamw" else optax.adam(lr)                # lunar_lander.py:48 / _hyper_params.py:41
    opt_state = optimizer.init(params)
    assert list(params.keys()) == MODULES and [tuple(params[m]["w"].shape) for m in MODULES] == [(D, 32), (32, 64), (64, 1), (64, A)]
    train_step = generate_train_step(optimizer, model)                                    # q_agent.py:112
    compute_q_targets = generate_q_target_comp(model, gamma, _Env(A))                     # q_agent.py:111
    compute_loss = generate_loss_computation(model)
    compute_action = action_computation(model)                                            # q_agent.py:110

    s, a, r, s2, d = synthetic_transitions(rng, N, D, A, done_p=done_p)
    out = dict(D=D, A=A, B=B, N=N, gamma=np.float64(gamma), lr=np.float64(lr), opt_kind=np.array(opt_kind), steps=steps,
               sync_at=np.array(sync_at, np.int64), states=s, actions=a, rewards=r, observations=s2, dones=d,
               theta_init=flat(params), target_init=flat(target_params))
    for t in range(steps):
        idx = rng.integers(0, N, B)
        if t == 1:
            idx[: B // 8] = idx[0]                           # duplicates: sampling is with replacement
        batch = (s[idx], a[idx], r[idx], s2[idx], d[idx])                                   # what sample_batch returns
        st, ac, rw, ob, dn = preprocessing(*batch)                                          # q_agent.py:154-158
        q_targets = compute_q_targets(params, target_params, st, ac, rw, ob, dn)           # q_agent.py:159-165
        loss = compute_loss(params, st, q_targets)
        grads = jax.grad(compute_loss)(params, st, q_targets)
        params, opt_state = train_step(params, opt_state, st, q_targets)                   # q_agent.py:166-169
        out[f"idx{t}"] = idx.astype(np.int64)
        out[f"q_targets{t}"] = np.asarray(q_targets, np.float32)
        out[f"loss{t}"] = np.float32(loss)
        out[f"grads{t}"] = flat(grads)
        out[f"theta{t}"] = flat(params)
        out[f"mu{t}"] = flat(opt_state[0].mu)
        out[f"nu{t}"] = flat(opt_state[0].nu)
        out[f"count{t}"] = np.int32(opt_state[0].count)
        if t in sync_at:                                                                    # q_agent.py:143-144
            target_params = params
        out[f"target{t}"] = flat(target_params)
    probe = rng.standard_normal((64, D)).astype(np.float32)
    out["probe_states"] = probe
    out["probe_actions"] = np.array([int(compute_action(params, probe[i:i + 1])) for i in range(64)], np.int32)   # q_agent.py:139
    out["probe_q"] = np.asarray(model.apply(params, probe), np.float32)
    np.savez_compressed(os.path.join(OUT, f"train_ref_{name}.npz"), **out)
    print(f"train_ref_{name}.npz: {steps} steps, final loss {float(out[f'loss{steps - 1}']):.6f}")


def main():
    # the shipped step-0 checkpoint (Test/lunar_lander/*.pickle), converted leaf by leaf by oracle/make_golden.py
    ck = np.load(os.path.join(OUT, "ref_checkpoint.npz"), allow_pickle=False)
    theta0 = {m: {k: ck[f"params|{m}|{k}"] for k in ("w", "b")} for m in MODULES}
    # Test/lunar_lander.py: D = 9, A = 4, B = 64, gamma .99, adamw(2e-4), theta_0 = the reference's own checkpoint
    run_case("lunar_lander", 9, 4, 64, 600, 0.99, "adamw", 2e-4, steps=6, sync_at=[2], theta0=theta0, seed=1)
    # Test/lunar_lander_hyper_params.py: adam(1e-4); sweep optimum batch 52, gamma 0.9028 (file comment :68-79)
    run_case("sweep", 9, 4, 52, 400, 0.9028, "adam", 1e-4, steps=5, sync_at=[1, 3], theta0=theta0, seed=2)
    # the sweep as the reference actually runs it: gamma frozen at 0.0; all-terminal batch (F5 quirk)
    run_case("gamma0_terminal", 9, 4, 38, 300, 0.0, "adam", 1e-3, steps=4, sync_at=[], theta0=theta0, seed=3, done_p=1.0)
    # BASELINE's synthetic shape: D = 8; two 64-row tiles
    run_case("d8_b70", 8, 4, 70, 500, 0.95, "adamw", 1e-3, steps=4, sync_at=[1], seed=4)


if __name__ == "__main__":
    main()
