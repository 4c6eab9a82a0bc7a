"""The numpy oracle (hand-derived backward, restated optax) against an independent derivation:
torch autograd + torch.optim.Adam/AdamW in float64.  This is the cross-check that stands in for the
missing jax/haiku/optax runtime (parity of the train-step half is otherwise unpinned)."""
import numpy as np
import pytest
import torch

from oracle import dqn_oracle as O
from oracle.replay_oracle import synthetic_transitions
from conftest import assert_close, golden_tree


def torch_forward(tp, x):
    l1, l2, lv, la = (tp[m] for m in O.MODULES)
    h1 = torch.relu(x @ l1["w"] + l1["b"])
    h2 = torch.relu(h1 @ l2["w"] + l2["b"])
    val = h2 @ lv["w"] + lv["b"]
    adv = h2 @ la["w"] + la["b"]
    return val + adv - adv.mean(dim=1, keepdim=True)


def to_torch(tree, requires_grad=False):
    return {m: {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=requires_grad)
                for k, v in tree[m].items()} for m in O.MODULES}


def torch_step(params, target_params, batch, gamma, loss_kind="huber"):
    """loss and grads of the reference's _step, by autograd (float64)."""
    s, a, r, s2, d = batch
    tp, tt = to_torch(params, True), to_torch(target_params)
    s_t, s2_t = torch.tensor(s, dtype=torch.float64), torch.tensor(s2, dtype=torch.float64)
    a_t = torch.tensor(a, dtype=torch.int64)
    r_t, d_t = torch.tensor(r, dtype=torch.float64), torch.tensor(d.astype(np.float64))
    with torch.no_grad():                                   # targets are inputs of train_step (F7)
        q = torch_forward(tp, s_t)
        nq = torch_forward(tp, s2_t)
        nqt = torch_forward(tt, s2_t)
        astar = nq.argmax(dim=1)
        rows = torch.arange(len(a))
        tv = r_t + (1.0 - d_t) * (gamma * nqt[rows, astar] - q[rows, a_t])
        targets = q + tv[:, None] * torch.nn.functional.one_hot(a_t, q.shape[1]).to(torch.float64)
    pred = torch_forward(tp, s_t)
    if loss_kind == "huber":
        loss = torch.nn.functional.huber_loss(pred, targets, reduction="none", delta=1.0).sum(dim=1).mean()
    else:                                                   # optax.l2_loss: 0.5 e^2
        loss = (0.5 * (pred - targets) ** 2).sum(dim=1).mean()
    loss.backward()
    return loss.item(), tp, targets.numpy(), astar.numpy()


@pytest.mark.parametrize("D,B,seed", [(9, 64, 0), (8, 38, 1), (9, 70, 2), (3, 1, 3)])
def test_loss_and_grads_match_autograd(D, B, seed):
    rng = np.random.default_rng(seed)
    A = 4
    params = O.init_params(rng, D, A, bias_std=0.05)
    target = O.tree_map(lambda x: (x + 0.02 * rng.standard_normal(x.shape)).astype(np.float32), params)
    batch = synthetic_transitions(rng, B, D, A, done_p=0.2)
    targets, parts = O.compute_q_targets(params, target, *O.preprocessing(*batch), 0.99, return_parts=True)
    loss, grads = O.loss_and_grads(params, batch[0], targets)
    tloss, tp, ttargets, tastar = torch_step(params, target, batch, 0.99)
    assert np.array_equal(parts["max_actions"], tastar)
    assert_close(targets, ttargets, what="targets")
    assert abs(float(loss) - tloss) <= 1e-5 * abs(tloss)
    for m in O.MODULES:
        for k in ("w", "b"):
            assert_close(grads[m][k], tp[m][k].grad.numpy().reshape(grads[m][k].shape), what=f"grad {m}/{k}")
    # both Huber branches must be exercised by the synthetic rewards (|TD| > 1 and < 1)
    e = np.abs(targets - parts["q"]).max(axis=1)
    if B >= 38:
        assert (e > 1).any() and (e < 1).any()


def test_l2_loss_extension_matches_autograd():
    """loss="l2" (0.5 e^2 in place of Huber; not in the reference): oracle vs torch autograd."""
    rng = np.random.default_rng(7)
    D, A, B = 8, 4, 64
    params = O.init_params(rng, D, A, bias_std=0.05)
    target = O.tree_map(lambda x: (x + 0.02 * rng.standard_normal(x.shape)).astype(np.float32), params)
    batch = synthetic_transitions(rng, B, D, A, done_p=0.2)
    targets = O.compute_q_targets(params, target, *O.preprocessing(*batch), 0.99)
    loss, grads = O.loss_and_grads(params, batch[0], targets, loss="l2")
    tloss, tp, _, _ = torch_step(params, target, batch, 0.99, loss_kind="l2")
    assert abs(float(loss) - tloss) <= 1e-5 * abs(tloss)
    assert float(loss) > float(O.loss_and_grads(params, batch[0], targets)[0])      # |TD| > 1 rows: quadratic > Huber
    for m in O.MODULES:
        for k in ("w", "b"):
            assert_close(grads[m][k], tp[m][k].grad.numpy().reshape(grads[m][k].shape), what=f"l2 grad {m}/{k}")


def test_polyak_extension():
    """theta^- := tau theta + (1 - tau) theta^- (optax.incremental_update); tau = 1 is the hard copy, tau = 0 a no-op."""
    rng = np.random.default_rng(8)
    p = O.init_params(rng, 9, 4, bias_std=0.05)
    t = O.tree_map(lambda x: (x + 0.1 * rng.standard_normal(x.shape)).astype(np.float32), p)
    for m in O.MODULES:
        assert np.array_equal(O.polyak(t, p, 1.0)[m]["w"], p[m]["w"]) and np.array_equal(O.polyak(t, p, 0.0)[m]["w"], t[m]["w"])
        got = O.polyak(t, p, 0.005)[m]["w"]
        want = 0.005 * p[m]["w"].astype(np.float64) + 0.995 * t[m]["w"].astype(np.float64)
        assert got.dtype == np.float32
        np.testing.assert_allclose(got, want, rtol=3e-7, atol=1e-9)


def test_terminal_state_quirk_F5():
    """done=1 rows: target for the taken action is q + r (TD error = -r), not r."""
    rng = np.random.default_rng(5)
    params = O.init_params(rng, 9, 4, bias_std=0.05)
    s, a, r, s2, d = synthetic_transitions(rng, 32, 9, 4, done_p=0.0)
    d[:] = True
    targets, parts = O.compute_q_targets(params, params, *O.preprocessing(s, a, r, s2, d), 0.99, return_parts=True)
    rows = np.arange(32)
    np.testing.assert_allclose(targets[rows, a] - parts["q"][rows, a], r, rtol=1e-5, atol=1e-6)
    mask = np.ones_like(targets, bool); mask[rows, a] = False
    assert np.array_equal(targets[mask], parts["q"][mask])          # untouched actions keep q exactly


def test_argmax_first_max_on_ties():
    params = O.tree_zeros_like(O.init_params(np.random.default_rng(0), 9, 4))
    assert O.compute_action(params, np.ones((1, 9), np.float32)) == 0      # all Q equal -> index 0
    s, a, r, s2, d = synthetic_transitions(np.random.default_rng(1), 8, 9, 4)
    _, parts = O.compute_q_targets(params, params, *O.preprocessing(s, a, r, s2, d), 0.9, return_parts=True)
    assert np.all(parts["max_actions"] == 0)


@pytest.mark.parametrize("kind,lr", [("adamw", 2e-4), ("adam", 1e-4)])
def test_ten_steps_match_torch_optimizer(kind, lr):
    """10 consecutive train steps incl. a hard target sync, oracle (fp32) vs torch AdamW/Adam (fp64)."""
    rng = np.random.default_rng(11)
    D, A, B = 9, 4, 64
    params = O.init_params(rng, D, A, bias_std=0.05)
    target = O.tree_copy(params)
    opt = O.OptSpec(kind, lr)
    state = O.init_opt_state(params)
    tp = to_torch(params, True)
    leaves = [tp[m][k] for m in O.MODULES for k in ("w", "b")]
    topt = (torch.optim.AdamW(leaves, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4) if kind == "adamw"
            else torch.optim.Adam(leaves, lr=lr, betas=(0.9, 0.999), eps=1e-8))
    ttarget = to_torch(target)
    for step in range(10):
        batch = synthetic_transitions(rng, B, D, A, done_p=0.2)
        params, state, parts = O.train_step(params, target, state, batch, 0.99, opt, return_parts=True)
        # torch side: same batch
        cur = {m: {k: tp[m][k].detach().numpy() for k in ("w", "b")} for m in O.MODULES}
        tgt = {m: {k: ttarget[m][k].numpy() for k in ("w", "b")} for m in O.MODULES}
        _, tp2, _, _ = torch_step(cur, tgt, batch, 0.99)
        topt.zero_grad()
        for m in O.MODULES:
            for k in ("w", "b"):
                tp[m][k].grad = tp2[m][k].grad.clone()
        topt.step()
        if step == 4:                                             # hard sync (q_agent.py:143-144)
            target = O.tree_copy(params)
            ttarget = {m: {k: tp[m][k].detach().clone() for k in ("w", "b")} for m in O.MODULES}
        for m in O.MODULES:
            for k in ("w", "b"):
                assert_close(params[m][k], tp[m][k].detach().numpy(), rtol=2e-5, what=f"step {step} {m}/{k}")
    assert int(state["count"]) == 10


def test_adam_first_step_identity():
    """t=1: m_hat = g, v_hat = g^2 -> update = -lr * g / (|g| + eps)."""
    rng = np.random.default_rng(3)
    params = O.init_params(rng, 9, 4)
    grads = O.tree_map(lambda x: rng.standard_normal(x.shape).astype(np.float32), params)
    new, st = O.adam_update(params, grads, O.init_opt_state(params), O.OptSpec("adam", 1e-3))
    for m in O.MODULES:
        g = grads[m]["w"].astype(np.float64)
        want = params[m]["w"] - 1e-3 * g / (np.abs(g) + 1e-8)
        np.testing.assert_allclose(new[m]["w"], want, rtol=2e-4, atol=1e-7)   # fp32 (1-b2^1) costs ~1e-4 rel
    assert int(st["count"]) == 1


def test_pow_f32_is_correctly_rounded():
    from fractions import Fraction
    for base in (0.9, 0.999):
        b = Fraction(float(np.float32(base)))
        for t in (1, 2, 3, 10, 100, 1000, 5000):
            exact = b ** t
            got = O.pow_f32(base, t)
            lo, hi = np.nextafter(got, np.float32(-np.inf)), np.nextafter(got, np.float32(np.inf))
            assert abs(Fraction(float(got)) - exact) <= min(abs(Fraction(float(lo)) - exact),
                                                            abs(Fraction(float(hi)) - exact))


def test_golden_checkpoint_is_usable_theta0(golden):
    ck = golden["ref_checkpoint"]
    params = golden_tree(ck, "params")
    assert [str(m) for m in ck["module_order"]] == list(O.MODULES)
    assert params[O.MODULES[0]]["w"].shape == (9, 32) and params[O.MODULES[3]]["w"].shape == (64, 4)
    assert all(v.dtype == np.float32 for m in params.values() for v in m.values())
    assert int(ck["count"]) == 0 and ck["count"].dtype == np.int32
    assert [str(x) for x in ck["opt_chain"]] == ["ScaleByAdamState", "EmptyState", "EmptyState"]   # adamw chain
    q = O.forward(params, np.zeros((2, 9), np.float32))
    assert q.shape == (2, 4) and np.allclose(q, 0)        # zero biases, zero input -> Q = 0
