"""``jax.numpy`` subset: float32 numpy ops, or the same op on torch tensors while ``jax.grad`` is tracing."""
import numpy as _np
import torch as _torch

ndarray = _np.ndarray
float32 = _np.float32


def _is_t(x):
    return isinstance(x, _torch.Tensor)


def asarray(x, dtype=None):
    return x if _is_t(x) else _np.asarray(x, dtype=dtype)


def array(x, dtype=None):
    if isinstance(x, (list, tuple)) and x and _is_t(x[0]):
        return _torch.stack(list(x))
    return _np.array(x, dtype=dtype if dtype is not None else (_np.float32 if _np.asarray(x).dtype.kind == "f" else None))


def mean(x, axis=None, keepdims=False):
    if _is_t(x):
        return _torch.mean(x) if axis is None else _torch.mean(x, dim=axis, keepdim=keepdims)
    return _np.mean(x, axis=axis, keepdims=keepdims, dtype=_np.float32)


def sum(x, axis=None, keepdims=False):  # noqa: A001
    if _is_t(x):
        return _torch.sum(x) if axis is None else _torch.sum(x, dim=axis, keepdim=keepdims)
    return _np.sum(x, axis=axis, keepdims=keepdims, dtype=_np.float32)


def dot(a, b):
    return a @ b


def abs(x):  # noqa: A001
    return _torch.abs(x) if _is_t(x) else _np.abs(x)


def minimum(a, b):
    if _is_t(a) or _is_t(b):
        a = a if _is_t(a) else _torch.tensor(a, dtype=_torch.float32)
        b = b if _is_t(b) else _torch.tensor(b, dtype=_torch.float32)
        return _torch.minimum(a, b)
    return _np.minimum(a, _np.float32(b) if _np.isscalar(b) else b)


def sqrt(x):
    return _torch.sqrt(x) if _is_t(x) else _np.sqrt(x)


def zeros_like(x):
    return _np.zeros_like(x)
