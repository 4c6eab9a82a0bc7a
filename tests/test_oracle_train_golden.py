"""Pins the train-step oracle (oracle/dqn_oracle.py) to golden vectors produced by EXECUTING the reference's own
``General/QLearning/q_learning_functions.py`` + ``LunarLander/dddqn.py`` (oracle/make_golden_train.py: the sources are
imported unmodified from /root/reference; only the absent third-party modules underneath them are shims).  Every step's
q-targets, loss, gradients, updated parameters, Adam moments / count, target network and the greedy actions must agree."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_close
from oracle import dqn_oracle as O

CASES = ["lunar_lander", "sweep", "gamma0_terminal", "d8_b70"]

# nu = b2 nu + (1 - b2) g^2 is QUADRATIC in the gradient: a gradient entry that meets the 1e-5 relative bar has its square
# within 2e-5 (first order), so nu's bar is twice the gradient's ("tolerances"); the Adam arithmetic
# itself is checked to 1e-6 against the kernel's own gradient in tests/test_gpu_train_step.py::test_adam_update_from_own_gradient
NU_RTOL = 2e-5


def unflat(flat, D, A):
    tree, o = {}, 0
    for name, (fi, fo) in zip(O.MODULES, [(D, 32), (32, 64), (64, 1), (64, A)]):
        tree[name] = {"w": flat[o:o + fi * fo].reshape(fi, fo).copy(), "b": flat[o + fi * fo:o + fi * fo + fo].copy()}
        o += fi * fo + fo
    assert o == flat.size
    return tree


def flat(tree):
    return np.concatenate([np.ravel(tree[m][k]) for m in O.MODULES for k in ("w", "b")])


def load(case):
    return np.load(os.path.join(GOLDEN, f"train_ref_{case}.npz"), allow_pickle=False)


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_reference_sources(case):
    g = load(case)
    D, A, B = int(g["D"]), int(g["A"]), int(g["B"])
    params, target = unflat(g["theta_init"], D, A), unflat(g["target_init"], D, A)
    opt = O.OptSpec(str(g["opt_kind"]), float(g["lr"]))
    opt_state = O.init_opt_state(params)
    s, a, r, s2, d = g["states"], g["actions"], g["rewards"], g["observations"], g["dones"]
    for t in range(int(g["steps"])):
        idx = g[f"idx{t}"]
        batch = (s[idx], a[idx], r[idx], s2[idx], d[idx])
        params, opt_state, parts = O.train_step(params, target, opt_state, batch, float(g["gamma"]), opt, return_parts=True)
        assert_close(parts["targets"], g[f"q_targets{t}"], what=f"{case} step {t} q_targets")
        assert abs(float(parts["loss"]) - float(g[f"loss{t}"])) <= 1e-5 * abs(float(g[f"loss{t}"])), f"{case} step {t} loss"
        assert_close(flat(parts["grads"]), g[f"grads{t}"], what=f"{case} step {t} grads")
        assert_close(flat(params), g[f"theta{t}"], what=f"{case} step {t} params")
        assert_close(flat(opt_state["mu"]), g[f"mu{t}"], what=f"{case} step {t} mu")
        assert_close(flat(opt_state["nu"]), g[f"nu{t}"], rtol=NU_RTOL, what=f"{case} step {t} nu")
        assert int(opt_state["count"]) == int(g[f"count{t}"]) == t + 1
        if t in set(g["sync_at"].tolist()):
            target = O.tree_copy(params)                         # q_agent.py:143-144
        assert_close(flat(target), g[f"target{t}"], what=f"{case} step {t} target")
    q = O.forward(params, g["probe_states"])
    assert_close(q, g["probe_q"], atol_scale=5e-6, what=f"{case} probe Q")     # Q = V + A - mean(A) cancels to ~0 in places
    srt = np.sort(g["probe_q"], axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-5
    got = np.array([O.compute_action(params, g["probe_states"][i:i + 1]) for i in range(64)])
    assert np.array_equal(got[clear], g["probe_actions"][clear])


def test_golden_exercises_the_quirks():
    g = load("gamma0_terminal")                                   # every transition terminal: target - q == reward
    assert g["dones"].all() and float(g["gamma"]) == 0.0
    g = load("lunar_lander")
    assert str(g["opt_kind"]) == "adamw" and int(g["D"]) == 9 and len(g["idx1"]) != len(set(g["idx1"].tolist()))
