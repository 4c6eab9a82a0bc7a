"""NumPy fp32 restatement of the reference's dueling double-DQN train step (oracle; test infra only).

Pinned by ``tests/golden/train_ref_*.npz`` (the reference's own ``q_learning_functions.py`` / ``dddqn.py`` executed over
``oracle/ref_shims``) -- see ``oracle/__init__.py``.  Every function cites the reference
lines it restates (paths relative to ``/root/reference``).  Third-party arithmetic that is not in
the reference tree (no version pinned by the reference; era: jax 0.3.x, dm-haiku 0.0.x,
optax 0.1.x) is restated from the libraries' published semantics:

* ``hk.Linear``      : ``x @ w + b``, ``w`` stored ``[in, out]``; default init
                       ``TruncatedNormal(stddev=1/sqrt(fan_in))`` (+-2 sigma), ``b = 0``.
* ``optax.huber_loss``: ``e = pred - tgt; q = min(|e|, delta); 0.5 q^2 + delta (|e| - q)``.
* ``optax.adam``     : ``scale_by_adam(b1=.9, b2=.999, eps=1e-8, eps_root=0)`` then ``scale(-lr)``.
* ``optax.adamw``    : same, with ``add_decayed_weights(1e-4)`` between the two (all leaves).
* ``scale_by_adam``  : ``count_inc = count + 1`` (int32, saturating);
                       ``mu = b1 mu + (1-b1) g``; ``nu = b2 nu + (1-b2) g^2``;
                       ``mu_hat = mu / (1 - b1**count_inc)``; ``nu_hat = nu / (1 - b2**count_inc)``;
                       ``u = mu_hat / (sqrt(nu_hat + eps_root) + eps)``.

Tree layout is the reference checkpoint's (``Test/lunar_lander/params.pickle``): module names
``model/~/linear`` (D->H1), ``linear_1`` (H1->H2), ``linear_2`` (value head, H2->1), ``linear_3``
(advantage head, H2->A), each ``{'w': f32[in,out], 'b': f32[out]}``.
"""
from collections import OrderedDict

import numpy as np

F32 = np.float32
MODULES = ("model/~/linear", "model/~/linear_1", "model/~/linear_2", "model/~/linear_3")


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
def init_params(rng, obs_dim, num_actions, hidden=(32, 64), bias_std=0.0):
    """Haiku-style initialisation of ``LunarLander/dddqn.py:17-22`` (Model.__init__).

    ``bias_std > 0`` randomises the biases (the reference initialises them to 0, which hides
    bias-gradient bugs; parity runs use non-zero biases).
    """
    h1, h2 = hidden
    shapes = [(obs_dim, h1), (h1, h2), (h2, 1), (h2, num_actions)]
    tree = OrderedDict()
    for name, (fi, fo) in zip(MODULES, shapes):
        std = 1.0 / np.sqrt(fi)
        w = rng.standard_normal((fi, fo))
        bad = np.abs(w) > 2.0
        while bad.any():                       # truncated normal at +-2 sigma, by rejection
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        b = rng.standard_normal(fo) * bias_std
        tree[name] = {"w": (w * std).astype(F32), "b": b.astype(F32)}
    return tree


def tree_map(fn, *trees):
    return OrderedDict((m, {k: fn(*[t[m][k] for t in trees]) for k in ("w", "b")}) for m in MODULES)


def tree_copy(tree):
    return tree_map(lambda x: np.array(x, dtype=F32, copy=True), tree)


def tree_zeros_like(tree):
    return tree_map(lambda x: np.zeros_like(x, dtype=F32), tree)


def tree_leaves(tree):
    """Leaves in the fixed order (w, b) per module -- used only for comparisons in tests."""
    return [tree[m][k] for m in MODULES for k in ("w", "b")]


# ------------------------------------------------------------------------------------------------
# A.1 forward  (LunarLander/dddqn.py:24-34)
# ------------------------------------------------------------------------------------------------
def forward(params, x, return_cache=False, return_pre=False):
    x = np.asarray(x, dtype=F32)
    l1, l2, lv, la = (params[m] for m in MODULES)
    z1 = x @ l1["w"] + l1["b"]                       # dddqn.py:25
    h1 = np.maximum(z1, F32(0))                      # :26  jax.nn.relu
    z2 = h1 @ l2["w"] + l2["b"]                      # :27
    h2 = np.maximum(z2, F32(0))                      # :28
    val = h2 @ lv["w"] + lv["b"]                     # :29  [B,1]
    adv = h2 @ la["w"] + la["b"]                     # :30  [B,A]
    q = val + adv - np.mean(adv, axis=1, keepdims=True, dtype=F32)   # :31
    q = q.astype(F32)
    if return_pre:
        return q, (x, h1, h2), (z1, z2)
    if return_cache:
        return q, (x, h1, h2)
    return q


# ------------------------------------------------------------------------------------------------
# A.7 action  (General/QLearning/q_learning_functions.py:67-73)
# ------------------------------------------------------------------------------------------------
def compute_action(params, state):
    """``argmax(network.apply(params, state))`` over the flattened [1,A] output, first max wins."""
    return int(np.argmax(forward(params, np.asarray(state, dtype=F32).reshape(1, -1))))


# ------------------------------------------------------------------------------------------------
# preprocessing  (q_learning_functions.py:76-85)
# ------------------------------------------------------------------------------------------------
def preprocessing(states, actions, rewards, observations, dones):
    return (np.asarray(states, dtype=F32), np.asarray(actions).astype(np.int32),
            np.asarray(rewards, dtype=F32), np.asarray(observations, dtype=F32),
            np.asarray(dones).astype(F32))          # :84  dones.astype(float32)


# ------------------------------------------------------------------------------------------------
# A.2 targets  (q_learning_functions.py:42-64)
# ------------------------------------------------------------------------------------------------
def compute_q_targets(params, target_params, states, actions, rewards, observations, dones, gamma,
                      return_parts=False):
    q = forward(params, states)                      # :52
    next_q = forward(params, observations)           # :53
    next_q_tm = forward(target_params, observations)  # :54
    max_actions = np.argmax(next_q, axis=1)          # :55  (first max on ties)
    g = F32(gamma)
    one = F32(1.0)
    b = q.shape[0]
    rows = np.arange(b)
    dones = np.asarray(dones, dtype=F32)
    rewards = np.asarray(rewards, dtype=F32)
    # :58  rewards[i] + (1.0 - done) * (gamma * next_q_tm[i, max_action] - q[i, action])
    target_val = rewards + (one - dones) * (g * next_q_tm[rows, max_actions] - q[rows, actions])
    target_val = target_val.astype(F32)
    targets = q.copy()                               # :59  q[i] + target_val * one_hot(action, A)
    targets[rows, actions] = q[rows, actions] + target_val
    if return_parts:
        return targets, dict(q=q, next_q=next_q, next_q_tm=next_q_tm,
                             max_actions=max_actions.astype(np.int32), target_val=target_val)
    return targets


# ------------------------------------------------------------------------------------------------
# A.3 loss  (q_learning_functions.py:31-39)
# ------------------------------------------------------------------------------------------------
def huber(pred, tgt, delta=1.0):
    e = (pred - tgt).astype(F32)
    ae = np.abs(e)
    quad = np.minimum(ae, F32(delta))
    return (F32(0.5) * quad * quad + F32(delta) * (ae - quad)).astype(F32), e


def l2_loss(pred, tgt):
    """``optax.l2_loss``: ``0.5 (pred - tgt)^2``.  NOT in the reference -- the self-specified "MSE" extension
    behind ``loss="l2"``; parity unpinned."""
    e = (pred - tgt).astype(F32)
    return (F32(0.5) * e * e).astype(F32), e


def compute_loss(params, states, q_targets, loss="huber"):
    pred = forward(params, states)                   # :35
    l, _ = huber(pred, q_targets) if loss == "huber" else l2_loss(pred, q_targets)
    return F32(np.mean(np.sum(l, axis=1, dtype=F32), axis=0, dtype=F32))   # :36


# ------------------------------------------------------------------------------------------------
# A.4 backward  (jax.grad(compute_loss), q_learning_functions.py:23)  -- hand-derived
# ------------------------------------------------------------------------------------------------
def relu_masks(z1, z2, device_h=None, tie_tol=0.0):
    """relu'(z) = [z > 0] for both hidden layers.  ``device_h = (h1, h2)`` (activations computed by the implementation
    under test) resolves NEAR-TIES: where ``|z| <= tie_tol * max|z|`` -- a pre-activation within fp32 round-off of zero,
    where the derivative of ReLU is discontinuous and either branch is a correctly rounded result -- the device's branch
    ``[h > 0]`` is taken; everywhere else the two must agree (asserted).  Returns (m1, m2, number of near-ties)."""
    out, ties = [], 0
    for i, z in enumerate((z1, z2)):
        m = z > 0
        if device_h is not None:
            near = np.abs(z) <= F32(tie_tol) * np.abs(z).max()
            dm = np.asarray(device_h[i]) > 0
            assert np.array_equal(dm[~near], m[~near]), "relu masks differ away from zero"
            m = np.where(near, dm, m)
            ties += int((near & (dm != (z > 0))).sum())
        out.append(m)
    return out[0], out[1], ties


def loss_and_grads(params, states, q_targets, loss="huber", device_h=None, tie_tol=0.0):
    """Loss and d(loss)/d(params) with the targets held constant (they are inputs)."""
    pred, (x, h1, h2), (z1, z2) = forward(params, states, return_pre=True)
    m1, m2, _ = relu_masks(z1, z2, device_h, tie_tol)
    b, a = pred.shape
    l, e = huber(pred, q_targets) if loss == "huber" else l2_loss(pred, q_targets)
    loss_value = F32(np.mean(np.sum(l, axis=1, dtype=F32), axis=0, dtype=F32))
    de = np.clip(e, F32(-1), F32(1)) if loss == "huber" else e        # d huber = clip(e,-1,1); d l2 = e
    dq = (de / F32(b)).astype(F32)                                    # mean over B
    loss = loss_value
    dval = np.sum(dq, axis=1, keepdims=True, dtype=F32)               # Q = V + A - mean(A)
    dadv = (dq - dval / F32(a)).astype(F32)
    l1, l2, lv, la = (params[m] for m in MODULES)
    grads = OrderedDict()
    grads[MODULES[2]] = {"w": (h2.T @ dval).astype(F32), "b": dval.sum(axis=0, dtype=F32)}
    grads[MODULES[3]] = {"w": (h2.T @ dadv).astype(F32), "b": dadv.sum(axis=0, dtype=F32)}
    dh2 = (dval @ lv["w"].T + dadv @ la["w"].T).astype(F32) * m2         # relu'(0) = 0
    dh2 = dh2.astype(F32)
    grads[MODULES[1]] = {"w": (h1.T @ dh2).astype(F32), "b": dh2.sum(axis=0, dtype=F32)}
    dh1 = ((dh2 @ l2["w"].T) * m1).astype(F32)
    grads[MODULES[0]] = {"w": (x.T @ dh1).astype(F32), "b": dh1.sum(axis=0, dtype=F32)}
    return loss, OrderedDict((m, grads[m]) for m in MODULES)


# ------------------------------------------------------------------------------------------------
# A.5 optimiser  (optimizer.update + optax.apply_updates, q_learning_functions.py:24-25)
# ------------------------------------------------------------------------------------------------
class OptSpec:
    """``optax.adam(lr)`` (``Test/lunar_lander_hyper_params.py:41``) or ``optax.adamw(lr)``
    (``Test/lunar_lander.py:48``; weight_decay defaults to 1e-4, no mask)."""

    def __init__(self, kind="adamw", lr=2e-4, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0, weight_decay=None):
        assert kind in ("adam", "adamw")
        self.kind, self.lr, self.b1, self.b2, self.eps, self.eps_root = kind, lr, b1, b2, eps, eps_root
        self.weight_decay = (1e-4 if kind == "adamw" else 0.0) if weight_decay is None else weight_decay


def pow_f32(base, t):
    """``decay ** count`` of optax's bias correction, as a correctly rounded fp32 value.

    jnp evaluates it as fp32 ``pow(f32(decay), f32(count))``; libm's powf is correctly rounded
    for essentially all arguments, so the double-precision pow rounded once to fp32 is the same
    number and is reproducible on the device (``pow`` in double, one thread per step).
    """
    return F32(np.power(np.float64(F32(base)), np.float64(int(t))))


def init_opt_state(params):
    """``optimizer.init(params)``: ``(count=0 int32, mu=0, nu=0)``."""
    return {"count": np.int32(0), "mu": tree_zeros_like(params), "nu": tree_zeros_like(params)}


def adam_update(params, grads, opt_state, opt):
    t = min(int(opt_state["count"]) + 1, np.iinfo(np.int32).max)   # safe_int32_increment
    b1, b2 = F32(opt.b1), F32(opt.b2)
    one = F32(1.0)
    c1 = one - pow_f32(opt.b1, t)
    c2 = one - pow_f32(opt.b2, t)
    eps, eps_root, lr, wd = F32(opt.eps), F32(opt.eps_root), F32(opt.lr), F32(opt.weight_decay)

    def upd_mu(g, m):
        return (b1 * m + (one - b1) * g).astype(F32)

    def upd_nu(g, v):
        return (b2 * v + (one - b2) * (g * g)).astype(F32)

    mu = tree_map(upd_mu, grads, opt_state["mu"])
    nu = tree_map(upd_nu, grads, opt_state["nu"])

    def new_param(p, m, v):
        u = (m / c1) / (np.sqrt(v / c2 + eps_root) + eps)
        if opt.weight_decay != 0.0:
            u = u + wd * p                              # add_decayed_weights
        return (p + (-lr) * u).astype(F32)             # scale(-lr); apply_updates

    new_params = tree_map(new_param, params, mu, nu)
    return new_params, {"count": np.int32(t), "mu": mu, "nu": nu}


# ------------------------------------------------------------------------------------------------
# Agent._step body  (General/QLearning/q_agent.py:154-169), given an already sampled batch
# ------------------------------------------------------------------------------------------------
def polyak(target_params, params, tau):
    """``optax.incremental_update(params, target_params, tau)``: ``tau * new + (1 - tau) * old`` leaf by leaf in fp32 (two
    rounded products, one rounded sum).  NOT in the reference (hard copy only) -- self-specified extension."""
    t = F32(tau)
    return tree_map(lambda new, old: (t * new + (F32(1.0) - t) * old).astype(F32), params, target_params)


def train_step(params, target_params, opt_state, batch, gamma, opt, return_parts=False, loss="huber", device_h=None, tie_tol=0.0):
    s, a, r, s2, d = preprocessing(*batch)                                   # q_agent.py:154-158
    targets, parts = compute_q_targets(params, target_params, s, a, r, s2, d, gamma,
                                       return_parts=True)                   # :159-165
    loss, grads = loss_and_grads(params, s, targets, loss, device_h, tie_tol)   # :166 (grad)
    new_params, new_opt_state = adam_update(params, grads, opt_state, opt)   # :166 (update/apply)
    if return_parts:
        parts.update(targets=targets, loss=loss, grads=grads)
        return new_params, new_opt_state, parts
    return new_params, new_opt_state
