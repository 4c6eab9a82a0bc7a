"""oracle/agent_oracle.py against the reference's own ``Agent._step()`` / ``ParamAgent.inject()`` entry points
(tests/golden/agent_step_ref.npz from oracle/make_golden_agent_step.py: real ``ReplayBuffer.add``, real numba
``sample_batch``, the reference's closures; the ring holds one transition repeated so the unseedable draw is moot)."""
import os

import numpy as np

from conftest import GOLDEN, assert_close
from oracle import dqn_oracle as O
from oracle.agent_oracle import OracleAgent


def unflat(flat, D=9, A=4):
    tree, o = {}, 0
    for name, (fi, fo) in zip(O.MODULES, [(D, 32), (32, 64), (64, 1), (64, A)]):
        tree[name] = {"w": flat[o:o + fi * fo].reshape(fi, fo).copy(), "b": flat[o + fi * fo:o + fi * fo + fo].copy()}
        o += fi * fo + fo
    return tree


def flat(tree):
    return np.concatenate([np.ravel(tree[m][k]) for m in O.MODULES for k in ("w", "b")])


def run(g, kind, lr, gamma, B, n_add, sync_after=None):
    theta = unflat(g["theta_init"])
    ag = OracleAgent(theta, O.init_opt_state(theta), O.OptSpec(kind, lr), int(g["N"]), 9, gamma, B, seed=0)
    for _ in range(n_add):
        ag.add(g["state"], int(g["action"]), float(g["reward"]), g["observation"], bool(g["done"]))
    out = []
    for t in range(3):
        ag.step()                                   # any index draw gives the same batch
        out.append(flat(ag.params))
        if t == sync_after:
            ag.update_target_model()
    return out


def test_agent_step_entry_point():
    g = np.load(os.path.join(GOLDEN, "agent_step_ref.npz"), allow_pickle=False)
    for t, th in enumerate(run(g, "adamw", 2e-4, 0.99, 64, 100, sync_after=1)):
        assert_close(th, g[f"agent_theta{t}"], what=f"Agent._step #{t + 1}")
    assert int(g["agent_count"]) == 3 and int(g["agent_buffer_size"]) == 100


def test_param_agent_inject_keeps_the_constructor_gamma():
    """SURVEY F12, on the reference's own ParamAgent: after inject(gamma=0.9028, ..., batch_size=52, ...) the step still
    discounts with the constructor's 0.0 (the closure captured it, q_agent.py:111); batch_size IS read live (:153)."""
    g = np.load(os.path.join(GOLDEN, "agent_step_ref.npz"), allow_pickle=False)
    assert float(g["pagent_gamma_attr"]) == 0.9028 and int(g["pagent_batch_size"]) == 52
    frozen = run(g, "adam", 1e-4, 0.0, 52, 60)
    live = run(g, "adam", 1e-4, 0.9028, 52, 60)
    for t in range(3):
        assert_close(frozen[t], g[f"pagent_theta{t}"], what=f"ParamAgent._step #{t + 1} (gamma frozen at 0)")
    # (three Adam steps move every weight by ~lr whatever the gradient's size, so the parameters barely tell the two
    #  discounts apart; the q-targets the agent's own closure computes do)
    theta = unflat(g["theta_init"])
    one = (g["state"][None, :], np.array([int(g["action"])]), np.array([float(g["reward"])], np.float32), g["observation"][None, :],
           np.array([0.0], np.float32))
    t_frozen = O.compute_q_targets(theta, theta, *one, 0.0)
    t_live = O.compute_q_targets(theta, theta, *one, 0.9028)
    assert_close(t_frozen, g["pagent_q_targets"], what="ParamAgent q-targets (gamma frozen at 0)")
    assert np.abs(t_live - g["pagent_q_targets"]).max() > 1e-3         # the injected discount is NOT what the reference used
