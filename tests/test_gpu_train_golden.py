"""The CUDA train step (through the C ABI) against golden vectors produced by EXECUTING the reference's own
``q_learning_functions.py`` / ``dddqn.py`` (oracle/make_golden_train.py; tests/golden/train_ref_*.npz): identical weights,
transitions and minibatch indices; q-targets, loss, gradients, parameters, Adam moments within 1e-5 relative in fp32
(north_star), Adam count and hard-synced target network exact, greedy actions bit-exact where the reference's own Q
margin is clear.  Both step kernels."""
import os

import numpy as np
import pytest

import dqn_b200
from conftest import GOLDEN, assert_close
from oracle import dqn_oracle as O

pytestmark = pytest.mark.gpu
CASES = ["lunar_lander", "sweep", "gamma0_terminal", "d8_b70"]


@pytest.fixture(autouse=True, params=["cta", "cluster"])
def step_kernel(request, monkeypatch):
    monkeypatch.setenv("DQN_B200_STEP_KERNEL", request.param)
    return request.param


@pytest.mark.parametrize("case", CASES)
def test_kernel_reproduces_the_reference_sources(case):
    g = np.load(os.path.join(GOLDEN, f"train_ref_{case}.npz"), allow_pickle=False)
    D, A, B, N = int(g["D"]), int(g["A"]), int(g["B"]), int(g["N"])
    kind, lr = str(g["opt_kind"]), float(g["lr"])
    eng = dqn_b200.DqnEngine(D, A, N, B, float(g["gamma"]), dqn_b200.adamw(lr) if kind == "adamw" else dqn_b200.adam(lr))
    eng.set_params_flat(g["theta_init"], 0, 0)
    eng.set_params_flat(g["target_init"], 0, 1)
    eng.store(g["states"], g["actions"], g["rewards"], g["observations"], g["dones"])
    for t in range(int(g["steps"])):
        got = eng.train_step_debug(indices=g[f"idx{t}"])
        assert np.array_equal(got["indices"], g[f"idx{t}"])
        assert_close(got["targets"], g[f"q_targets{t}"], what=f"{case} step {t} q_targets")
        assert abs(float(got["loss"]) - float(g[f"loss{t}"])) <= 1e-5 * abs(float(g[f"loss{t}"])), f"{case} step {t} loss"
        assert_close(got["grads_flat"], g[f"grads{t}"], what=f"{case} step {t} grads")
        assert_close(eng.get_params_flat(0, 0), g[f"theta{t}"], what=f"{case} step {t} params")
        cnt, mu, nu = eng.get_opt_state(0)
        flat = lambda tree: np.concatenate([np.ravel(tree[m][k]) for m in O.MODULES for k in ("w", "b")])
        assert int(cnt) == int(g[f"count{t}"])
        assert_close(flat(mu), g[f"mu{t}"], what=f"{case} step {t} mu")
        assert_close(flat(nu), g[f"nu{t}"], rtol=3e-5, what=f"{case} step {t} nu")       # quadratic in the gradient
        if t in set(g["sync_at"].tolist()):
            eng.sync_target()
            assert np.array_equal(eng.get_params_flat(0, 1), eng.get_params_flat(0, 0))
        assert_close(eng.get_params_flat(0, 1), g[f"target{t}"], what=f"{case} step {t} target")
    srt = np.sort(g["probe_q"], axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-4
    got = np.array([eng.act(s) for s in g["probe_states"]])
    assert clear.sum() > 50 and np.array_equal(got[clear], g["probe_actions"][clear])
