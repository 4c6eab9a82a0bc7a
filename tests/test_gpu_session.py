"""Session mode (dqn_set_session): one resident launch of the cluster train-step kernel serves add+step / act / hard
sync from commands in mapped host memory.  It must be indistinguishable from the launch-per-call path: same ring, same
parameters and moments bit for bit, same losses and actions -- across idle time-outs of the resident kernel and
across calls that are not served by it (which end the session and start a new one later)."""
import time

import numpy as np
import pytest

import dqn_b200
from oracle import dqn_oracle as O
from oracle.replay_oracle import synthetic_transitions

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
_lib = dqn_b200.pkg._lib


def pair(D=9, A=4, N=300, B=32, seed=5):
    """(launch-per-call engine, session engine) with identical state."""
    theta = O.init_params(np.random.default_rng(seed), D, A, bias_std=0.05)
    out = []
    for session in (False, True):
        e = dqn_b200.DqnEngine(D, A, N, B, 0.99, dqn_b200.adamw(1e-3), seed=seed + 1, step_kernel="cluster", session=session)
        e.set_params(theta, 0, 0)
        e.set_params(theta, 0, 1)
        out.append(e)
    return out


def store_step(e, data, loss):
    s, a, r, s2, d = data
    d8 = np.ascontiguousarray(d, dtype=np.bool_)
    _lib.check(e.lib.dqn_store_train_step(e.h, 0, len(a), _lib.ptr(s), _lib.ptr(a), _lib.ptr(r), _lib.ptr(s2), _lib.ptr(d8), 1,
                                          None if loss is None else _lib.ptr(loss)))


def same_state(ref, ses):
    assert np.array_equal(ref.get_params_flat(0, 0), ses.get_params_flat(0, 0))
    assert np.array_equal(ref.get_params_flat(0, 1), ses.get_params_flat(0, 1))
    (c0, m0, v0), (c1, m1, v1) = ref.get_opt_state(), ses.get_opt_state()
    assert c0 == c1
    for m in O.MODULES:
        for k in ("w", "b"):
            assert np.array_equal(m0[m][k], m1[m][k]) and np.array_equal(v0[m][k], v1[m][k])
    assert ref.buffer_state() == ses.buffer_state()
    for x, y in zip(ref.buffer_export(), ses.buffer_export()):
        assert np.array_equal(x, y)


def test_session_equals_launch_per_call_bitwise():
    ref, ses = pair()
    rng = np.random.default_rng(0)
    l0, l1 = np.zeros(1, np.float32), np.zeros(1, np.float32)
    for it in range(120):
        n = int(rng.integers(1, 9))                                   # wraps the 300-slot ring a few times
        data = synthetic_transitions(rng, n, 9, 4, done_p=0.2)
        with_loss = it % 3 != 0
        store_step(ref, data, l0 if with_loss else None)
        store_step(ses, data, l1 if with_loss else None)
        if with_loss:
            assert l0[0] == l1[0], f"loss at iteration {it}"
        else:
            assert ref.last_loss() == ses.last_loss()
        st = rng.standard_normal(9).astype(np.float32)
        assert ref.act(st) == ses.act(st)                             # greedy action from the resident weights
        if it % 25 == 24:
            ref.sync_target()
            ses.sync_target()
        if it == 60:
            same_state(ref, ses)                                      # reads end the session; the next call starts a new one
    same_state(ref, ses)
    assert ref.train_step_count() == ses.train_step_count() == 120


def test_session_survives_idle_timeouts_and_unserved_calls():
    ref, ses = pair(seed=9)
    rng = np.random.default_rng(1)
    l0, l1 = np.zeros(1, np.float32), np.zeros(1, np.float32)
    for it in range(12):
        data = synthetic_transitions(rng, 4, 9, 4, done_p=0.2)
        store_step(ref, data, l0)
        store_step(ses, data, None if it % 2 else l1)                 # odd iterations leave the command in flight ...
        if it % 4 == 1:
            time.sleep(0.08)                                          # ... across the resident kernel's ~30 ms idle time-out
        if it % 4 == 2:
            time.sleep(0.015)                                         # past the host's 10 ms lease: kernel retired and relaunched
        if it % 2 == 0:
            assert l0[0] == l1[0]
        if it == 7:
            ses.train_steps(3)                                        # a call the session does not serve (K = 3)
            ref.train_steps(3)
    assert ref.last_loss() == ses.last_loss()
    same_state(ref, ses)
    ses.set_session(False)
    store_step(ref, synthetic_transitions(rng, 2, 9, 4), l0)
    store_step(ses, synthetic_transitions(np.random.default_rng(1), 0, 9, 4), None)   # n = 0 is allowed
    with pytest.raises(dqn_b200.DqnError):
        dqn_b200.DqnEngine(9, 4, 100, 8, 0.9, dqn_b200.adam(1e-3), n_agents=2, session=True)   # single-agent handles only


def test_two_commands_in_flight_lagged_loss():
    """The pipelined env loop: publish step i + 1 while step i may still run, then read step i's loss
    (dqn_get_loss_lagged).  Losses, parameters, moments and ring must equal the launch-per-call engine bit for bit; idle
    time-outs with two commands pending and unserved calls in between must not lose a command or a loss."""
    ref, ses = pair(seed=31)
    rng = np.random.default_rng(7)
    ref_losses, ses_losses = [], []
    l0 = np.zeros(1, np.float32)
    for it in range(200):
        n = int(rng.integers(1, 7))
        data = synthetic_transitions(rng, n, 9, 4, done_p=0.2)
        store_step(ref, data, l0)
        ref_losses.append(float(l0[0]))
        store_step(ses, data, None)                                   # published; step it - 1 may still be running
        if it:
            ses_losses.append(ses.lagged_loss(0, 1))                  # loss of step it - 1
        if it % 37 == 36:
            time.sleep(0.08)                                          # resident kernel leaves with a command answered but not collected
        if it == 90:
            st = rng.standard_normal(9).astype(np.float32)
            assert ref.act(st) == ses.act(st)                         # a command that drains both slots
        if it == 120:
            ref.sync_target(); ses.sync_target()
        if it == 150:
            ses.train_steps(2); ref.train_steps(2)                    # not served by the session: ends it, lagged loss falls back to the ring
            ref_losses.append(None); ref_losses.append(float(ref.last_loss()))
            ses_losses.append(None); ses_losses.append(ses.lagged_loss(0, 1))
            ses_losses.append(float(ses.last_loss()))
    ses_losses.append(float(ses.last_loss()))
    # drop the two bookkeeping entries around the K = 2 launch (its first step's loss is only in the ring)
    ref_seq = [x for x in ref_losses if x is not None]
    ses_seq = [x for x in ses_losses if x is not None]
    assert ref.lagged_loss(0, 1) == ses.lagged_loss(0, 1)
    assert ref.train_step_count() == ses.train_step_count() == 202
    same_state(ref, ses)
    assert len(ref_seq) >= 200
    assert ses_seq[:150] == ref_seq[:150]
    assert ses_seq[-40:] == ref_seq[-40:]


def test_session_large_payloads():
    """D = 16 (160-byte records) and up to 16 adds per step: 640 payload units, i.e. several read rounds beyond the 128
    units that ride along with the doorbell poll; A = 7, B = 130 (three tiles) for good measure."""
    ref, ses = pair(D=16, A=7, N=200, B=130, seed=17)
    rng = np.random.default_rng(3)
    l0, l1 = np.zeros(1, np.float32), np.zeros(1, np.float32)
    for it, n in enumerate([16, 16, 1, 9, 16, 5, 16, 13, 16, 2]):
        data = synthetic_transitions(rng, n, 16, 7, done_p=0.3)
        store_step(ref, data, l0)
        store_step(ses, data, l1)
        assert l0[0] == l1[0], f"iteration {it}"
        st = rng.standard_normal(16).astype(np.float32)
        assert ref.act(st) == ses.act(st)
    same_state(ref, ses)


def test_command_sent_to_a_timed_out_kernel_is_resent():
    """Without the host-side lease (dqn_set_session(h, 2)) a command can reach a kernel that has already left on its idle
    time-out: the host notices the drained stream and serves the same command from a fresh launch."""
    ref, ses = pair(seed=21)
    ses.set_session(2)
    rng = np.random.default_rng(2)
    l0, l1 = np.zeros(1, np.float32), np.zeros(1, np.float32)
    for it in range(6):
        data = synthetic_transitions(rng, 3, 9, 4, done_p=0.2)
        store_step(ref, data, l0)
        store_step(ses, data, l1)
        assert l0[0] == l1[0]
        st = rng.standard_normal(9).astype(np.float32)
        time.sleep(0.07)                                              # the resident kernel leaves (~30 ms idle)
        assert ref.act(st) == ses.act(st)                             # ... and this command finds it gone
        time.sleep(0.07)
    same_state(ref, ses)


def test_agent_dropin_with_session():
    """Agent(session=True): _policy / add / _step / _update_target_model through the resident kernel == the plain Agent."""
    import asyncio
    theta = O.init_params(np.random.default_rng(3), 9, 4, bias_std=0.05)
    agents = []
    for session in (False, True):
        opt = dqn_b200.adamw(2e-4)
        agents.append(dqn_b200.Agent(network=dqn_b200.Model(4), params=theta, optimizer=opt, opt_state=opt.init(theta), env=None,
                                     buffer_size=500, obs_shape=(500, 9), ac_shape=(500,), gamma=0.99, epsilon=0.0,
                                     epsilon_decay_rate=0.99, min_epsilon=0.0, max_episodes=10, max_steps=1500, training_start=50,
                                     batch_size=32, train_frequency=4, back_up_frequency=50, replace_frequency=20,
                                     reward_to_reach=230.0, num_actions=4, saving_directory="/tmp/dqn_b200_test_session",
                                     seed=13, session=session))
    rng = np.random.default_rng(4)
    s, a, r, s2, d = synthetic_transitions(rng, 200, 9, 4, done_p=0.1)
    for i in range(200):
        acts = [ag._policy(s[i:i + 1]) for ag in agents]              # epsilon = 0: greedy branch
        assert acts[0] == acts[1]
        for ag in agents:
            ag._replay_buffer.add(s[i], int(a[i]), float(r[i]), s2[i], bool(d[i]))
            if i >= 50 and i % 4 == 0:
                ag._step()
            if i % 70 == 69:
                asyncio.run(ag._update_target_model())
    p0, p1 = agents[0]._params, agents[1]._params
    for m in O.MODULES:
        assert np.array_equal(p0[m]["w"], p1[m]["w"]) and np.array_equal(p0[m]["b"], p1[m]["b"])
    t0, t1 = agents[0]._target_params, agents[1]._target_params
    for m in O.MODULES:
        assert np.array_equal(t0[m]["w"], t1[m]["w"])
