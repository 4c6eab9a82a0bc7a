"""Generate ``tests/golden/agent_step_ref.npz`` by calling the reference's own ``Agent._step()`` and
``ParamAgent.inject()`` + ``_step()`` (build container only; needs /root/reference and numba).

``Agent._step`` (``General/QLearning/q_agent.py:146-169``) draws its minibatch with numba's unseedable RNG
(``General/Base/replay_buffer.py:77``), so the replay ring is filled with ONE transition repeated: every
possible draw then yields the same batch and the entry point itself -- real ``ReplayBuffer.add``, real numba
``sample_batch``, the reference's closures over ``oracle/ref_shims`` -- becomes reproducible end to end.

Two agents are recorded:
  * ``Agent`` as ``Test/lunar_lander.py`` builds it (gamma .99, adamw 2e-4, batch 64): parameters after 1, 2, 3 ``_step()``s,
    with an ``_update_target_model()`` after the second;
  * ``ParamAgent`` as ``Test/lunar_lander_hyper_params.py`` builds it (constructor gamma 0.) after
    ``inject(gamma=0.9028, ..., batch_size=52, ...)``: the discount its ``_step()`` uses is the one captured by the
    closure at construction (0.0), not the injected one -- the parameters after the steps show which.
"""
import asyncio
import os
import sys
import tempfile

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shims"))

import haiku as hk                                                   # noqa: E402  (shim)
import optax                                                         # noqa: E402  (shim)
from General.QLearning.hyperparameter_optimization import ParamAgent  # noqa: E402  (the reference's own code)
from General.QLearning.q_agent import Agent                          # noqa: E402  (the reference's own code)
from LunarLander.dddqn import Model                                  # noqa: E402  (the reference's own code)

D, A, N = 9, 4, 256
MODULES = ["model/~/linear", "model/~/linear_1", "model/~/linear_2", "model/~/linear_3"]


class _Space:
    n = A


class _Env:
    action_space = _Space()


def flat(tree):
    return np.concatenate([np.ravel(np.asarray(tree[m][k], np.float32)) for m in MODULES for k in ("w", "b")])


def main():
    ck = np.load(os.path.join(OUT, "ref_checkpoint.npz"), allow_pickle=False)
    rng = np.random.default_rng(5)
    params = {m: {k: np.array(ck[f"params|{m}|{k}"], np.float32) for k in ("w", "b")} for m in MODULES}
    for m in MODULES:
        params[m]["b"] = (0.05 * rng.standard_normal(params[m]["b"].shape)).astype(np.float32)
    model = hk.without_apply_rng(hk.transform(lambda *args: Model(A)(*args)))
    state, obs = rng.standard_normal(D).astype(np.float32), rng.standard_normal(D).astype(np.float32)
    action, reward, done = 2, 1.75, False
    out = dict(theta_init=flat(params), state=state, observation=obs, action=np.int64(action), reward=np.float32(reward), done=np.bool_(done), N=N)

    # ---- Agent, Test/lunar_lander.py wiring (:39-77) ----
    opt = optax.adamw(0.0002)
    agent = Agent(network=model, params=params, optimizer=opt, opt_state=opt.init(params), env=_Env(), buffer_size=N,
                  obs_shape=(N, D), ac_shape=(N,), gamma=0.99, epsilon=1.0, epsilon_decay_rate=0.99, min_epsilon=0.15,
                  max_episodes=10000, max_steps=1500, training_start=250, batch_size=64, train_frequency=4, back_up_frequency=50,
                  replace_frequency=20, reward_to_reach=230.0, num_actions=A, saving_directory=tempfile.mkdtemp(), monitoring=False)
    for _ in range(100):                                             # q_agent.py:182, the same transition every time
        agent._replay_buffer.add(state, action, reward, obs, done)
    for t in range(3):
        agent._step()                                                # q_agent.py:146-169, numba sampler and all
        out[f"agent_theta{t}"] = flat(agent._params)
        if t == 1:
            asyncio.run(agent._update_target_model())                # q_agent.py:143-144
    out["agent_count"] = np.int32(agent._opt_state[0].count)
    out["agent_buffer_size"] = np.int64(agent._replay_buffer.size)

    # ---- ParamAgent, Test/lunar_lander_hyper_params.py wiring (:32-63) + inject (hyperparameter_optimization.py:76-91) ----
    opt = optax.adam(0.0001)
    pagent = ParamAgent(network=model, params=params, optimizer=opt, opt_state=opt.init(params), env=_Env(), buffer_size=N,
                        obs_shape=(N, D), ac_shape=(N,), max_episodes=10000, max_steps=1500, training_start=500,
                        back_up_frequency=50, reward_to_reach=240.0, num_actions=A, saving_directory=tempfile.mkdtemp())
    pagent.inject(0.9028, 0.979, 0.9873, 0.1469, 25, 52, 7)          # the optimum recorded in the sweep script (:68-79)
    for _ in range(60):
        pagent._replay_buffer.add(state, action, reward, obs, done)
    # the closure the agent's _step() calls (q_agent.py:159), on the one-transition batch: which discount is in it?
    one = (state[None, :], np.array([action]), np.array([reward], np.float32), obs[None, :], np.array([0.0], np.float32))
    out["pagent_q_targets"] = np.asarray(pagent._compute_q_targets(pagent._params, pagent._target_params, *one), np.float32)
    for t in range(3):
        pagent._step()
        out[f"pagent_theta{t}"] = flat(pagent._params)
    out["pagent_batch_size"] = np.int64(pagent._batch_size)
    out["pagent_gamma_attr"] = np.float64(pagent._gamma)
    np.savez_compressed(os.path.join(OUT, "agent_step_ref.npz"), **out)
    print("agent_step_ref.npz written; |theta3 - theta0| max (Agent) = %.3e" % np.abs(out["agent_theta2"] - out["theta_init"]).max())


if __name__ == "__main__":
    main()
