"""Shim of the optax pieces the reference uses: ``adam`` / ``adamw`` (a ``chain`` of ``scale_by_adam``,
``add_decayed_weights`` and ``scale``), ``huber_loss``, ``apply_updates`` -- optax 0.1.x semantics, float32."""
from collections import namedtuple

import numpy as _np

import jax.numpy as jnp

ScaleByAdamState = namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
EmptyState = namedtuple("EmptyState", [])
GradientTransformation = namedtuple("GradientTransformation", ["init", "update"])


def _tree_map(f, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return type(t0)((k, _tree_map(f, *[t[k] for t in trees])) for k in t0)
    return f(*trees)


def huber_loss(predictions, targets=None, delta=1.0):
    errors = (predictions - targets) if targets is not None else predictions
    abs_errors = jnp.abs(errors)
    quadratic = jnp.minimum(abs_errors, delta)
    linear = abs_errors - quadratic
    return 0.5 * quadratic ** 2 + delta * linear


def _bias_correction(moment, decay, count):
    # optax: moment / (1 - decay ** count).  `decay` is a weakly typed Python float and `count` an int32 array, so jax
    # evaluates the power and the subtraction in float32.
    c = _np.float32(1) - _np.power(_np.float32(decay), _np.float32(count), dtype=_np.float32)
    return _tree_map(lambda t: (t / c).astype(_np.float32), moment)


def scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        z = lambda p: _np.zeros_like(_np.asarray(p, dtype=_np.float32))
        return ScaleByAdamState(count=_np.zeros([], _np.int32), mu=_tree_map(z, params), nu=_tree_map(z, params))

    def update(updates, state, params=None):
        f = _np.float32
        mu = _tree_map(lambda g, m: (f(1 - b1) * g + f(b1) * m).astype(_np.float32), updates, state.mu)
        nu = _tree_map(lambda g, v: (f(1 - b2) * (g * g) + f(b2) * v).astype(_np.float32), updates, state.nu)
        count = _np.int32(min(int(state.count) + 1, 2**31 - 1))                     # safe_int32_increment
        mu_hat, nu_hat = _bias_correction(mu, b1, count), _bias_correction(nu, b2, count)
        out = _tree_map(lambda m, v: (m / (_np.sqrt(v + f(eps_root)) + f(eps))).astype(_np.float32), mu_hat, nu_hat)
        return out, ScaleByAdamState(count=count, mu=mu, nu=nu)

    return GradientTransformation(init, update)


def add_decayed_weights(weight_decay=0.0):
    def update(updates, state, params):
        return _tree_map(lambda g, p: (g + _np.float32(weight_decay) * p).astype(_np.float32), updates, params), state

    return GradientTransformation(lambda params: EmptyState(), update)


def scale(step_size):
    def update(updates, state, params=None):
        return _tree_map(lambda g: (_np.float32(step_size) * g).astype(_np.float32), updates), state

    return GradientTransformation(lambda params: EmptyState(), update)


def chain(*transforms):
    def init(params):
        return tuple(t.init(params) for t in transforms)

    def update(updates, state, params=None):
        new_state = []
        for t, s in zip(transforms, state):
            updates, s = t.update(updates, s, params)
            new_state.append(s)
        return updates, tuple(new_state)

    return GradientTransformation(init, update)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    return chain(scale_by_adam(b1, b2, eps, eps_root), scale(-learning_rate))


def adamw(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0, weight_decay=1e-4):
    return chain(scale_by_adam(b1, b2, eps, eps_root), add_decayed_weights(weight_decay), scale(-learning_rate))


def apply_updates(params, updates):
    return _tree_map(lambda p, u: (_np.asarray(p, dtype=_np.float32) + u).astype(_np.float32), params, updates)
