"""Light stand-ins for the reference's ``network`` / ``optimizer`` constructor arguments.

The reference passes a ``hk.Transformed`` (``Test/lunar_lander.py:47``) and an optax
``GradientTransformation`` (``:48``) into ``Agent``.  Neither library exists on the B200 path, so the
two arguments become plain spec objects carrying exactly what those objects determined:

* ``Model(num_actions)``  -- the dueling MLP of ``LunarLander/dddqn.py:11-34``
  (``D -> 32 -> 64 -> {V:1, A:num_actions}``); ``model.init(rng, test_input)`` gives a parameter tree
  with the checkpoint's module names and ``[in,out]`` weight layout.
* ``adam(lr)`` / ``adamw(lr)`` -- optax's defaults (b1=.9, b2=.999, eps=1e-8, eps_root=0, adamw
  weight_decay=1e-4, no mask); ``optimizer.init(params)`` gives ``(ScaleByAdamState(count, mu, nu),
  ...)`` in the structure of ``Test/lunar_lander/opt_state.pickle``.

Also here: the tree <-> flat conversion for the C ABI's parameter layout (``include/dqn_b200.h``).
"""
from collections import OrderedDict, namedtuple

import numpy as np

MODULES = ("model/~/linear", "model/~/linear_1", "model/~/linear_2", "model/~/linear_3")
HIDDEN = (32, 64)

ScaleByAdamState = namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
EmptyState = namedtuple("EmptyState", [])


def layer_shapes(obs_dim, num_actions, hidden=HIDDEN):
    h1, h2 = hidden
    return [(obs_dim, h1), (h1, h2), (h2, 1), (h2, num_actions)]


def param_count(obs_dim, num_actions, hidden=HIDDEN):
    return sum(fi * fo + fo for fi, fo in layer_shapes(obs_dim, num_actions, hidden))


def flatten_tree(tree, obs_dim, num_actions, hidden=HIDDEN):
    """Parameter tree -> f32[P] in the C ABI's flat order (w then b, module by module)."""
    parts = []
    for name, (fi, fo) in zip(MODULES, layer_shapes(obs_dim, num_actions, hidden)):
        w = np.asarray(tree[name]["w"], dtype=np.float32)
        b = np.asarray(tree[name]["b"], dtype=np.float32)
        if w.shape != (fi, fo) or b.shape != (fo,):
            raise ValueError(f"{name}: expected w{(fi, fo)} b{(fo,)}, got w{w.shape} b{b.shape}")
        parts += [w.reshape(-1), b]
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def unflatten_tree(flat, obs_dim, num_actions, hidden=HIDDEN):
    flat = np.asarray(flat, dtype=np.float32)
    tree, o = OrderedDict(), 0
    for name, (fi, fo) in zip(MODULES, layer_shapes(obs_dim, num_actions, hidden)):
        w = flat[o:o + fi * fo].reshape(fi, fo).copy(); o += fi * fo
        b = flat[o:o + fo].copy(); o += fo
        tree[name] = {"w": w, "b": b}
    if o != flat.size:
        raise ValueError(f"flat parameter vector has {flat.size} entries, expected {o}")
    return tree


class Model:
    """Spec of ``LunarLander/dddqn.py::Model`` wrapped in ``hk.without_apply_rng(hk.transform(...))``."""

    def __init__(self, num_actions, hidden=HIDDEN):
        if tuple(hidden) != HIDDEN:
            raise ValueError("the fused B200 path implements the reference widths (32, 64) only")
        self.num_actions = int(num_actions)
        self.hidden = HIDDEN

    def init(self, rng, test_input):
        """``model.init(rng, test_input)`` (``Test/lunar_lander.py:50``): haiku Linear defaults --
        ``w ~ TruncatedNormal(0, 1/sqrt(fan_in))`` truncated at +-2 sigma, ``b = 0``.  ``rng`` is an
        int seed or a ``numpy.random.Generator`` (jax PRNG keys do not exist here)."""
        gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        obs_dim = int(np.asarray(test_input).shape[-1])
        tree = OrderedDict()
        for name, (fi, fo) in zip(MODULES, layer_shapes(obs_dim, self.num_actions)):
            w = gen.standard_normal((fi, fo))
            bad = np.abs(w) > 2.0
            while bad.any():
                w[bad] = gen.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            tree[name] = {"w": (w / np.sqrt(fi)).astype(np.float32), "b": np.zeros(fo, np.float32)}
        return tree


class Optimizer:
    """Spec of ``optax.adam`` / ``optax.adamw`` (``q_learning_functions.py:24``)."""

    def __init__(self, kind, learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0, weight_decay=0.0):
        if kind not in ("adam", "adamw"):
            raise ValueError("optimizer kind must be 'adam' or 'adamw'")
        self.kind, self.learning_rate = kind, float(learning_rate)
        self.b1, self.b2, self.eps, self.eps_root = float(b1), float(b2), float(eps), float(eps_root)
        self.weight_decay = float(weight_decay) if kind == "adamw" else 0.0

    def init(self, params):
        zeros = lambda: OrderedDict((m, {"b": np.zeros_like(params[m]["b"], dtype=np.float32),
                                         "w": np.zeros_like(params[m]["w"], dtype=np.float32)})
                                    for m in params)
        adam_state = ScaleByAdamState(np.zeros((), np.int32), zeros(), zeros())
        if self.kind == "adamw":      # chain(scale_by_adam, add_decayed_weights, scale)
            return (adam_state, EmptyState(), EmptyState())
        return (adam_state, EmptyState())   # chain(scale_by_adam, scale)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    return Optimizer("adam", learning_rate, b1, b2, eps, eps_root)


def adamw(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0, weight_decay=1e-4):
    return Optimizer("adamw", learning_rate, b1, b2, eps, eps_root, weight_decay)


def split_opt_state(opt_state):
    """``(count, mu_tree, nu_tree)`` from an optax-style chain state (tuple whose first element is the
    ScaleByAdamState triple) or from a bare triple."""
    first = opt_state[0]
    if isinstance(first, (tuple, list)) and len(first) == 3:
        count, mu, nu = first
    else:
        count, mu, nu = opt_state
    return int(np.asarray(count)), mu, nu
