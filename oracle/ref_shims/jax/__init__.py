"""Shim of the parts of ``jax`` the reference's hot path uses (see ../README.md).  Test infrastructure only."""
import numpy as _np
import torch as _torch

from . import nn, numpy  # noqa: F401  (jax.nn, jax.numpy)


def jit(fun=None, **_kw):
    """``jax.jit`` compiles; it does not change what a function computes."""
    if fun is None:
        return lambda f: f
    return fun


def _map(tree, f):
    if isinstance(tree, dict):
        return type(tree)((k, _map(v, f)) for k, v in tree.items())
    if isinstance(tree, (tuple, list)):
        return type(tree)(_map(v, f) for v in tree)
    return f(tree)


def grad(fun):
    """``jax.grad(fun)``: gradient of a scalar-valued function w.r.t. its first argument (a pytree of arrays)."""

    def grad_fun(params, *args):
        leaves = []

        def to_t(x):
            t = _torch.tensor(_np.asarray(x, dtype=_np.float32), requires_grad=True)
            leaves.append(t)
            return t

        p = _map(params, to_t)
        targs = [_torch.as_tensor(_np.asarray(a, dtype=_np.float32)) if isinstance(a, _np.ndarray) else a for a in args]
        out = fun(p, *targs)
        out.backward()
        it = iter(leaves)
        return _map(params, lambda _x: next(it).grad.detach().numpy().astype(_np.float32))

    return grad_fun
