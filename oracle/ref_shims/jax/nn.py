"""``jax.nn`` subset."""
import numpy as _np
import torch as _torch


def relu(x):
    """max(x, 0); jax defines the gradient at exactly 0 as 0 (so does torch.relu)."""
    return _torch.relu(x) if isinstance(x, _torch.Tensor) else _np.maximum(x, _np.float32(0))


def one_hot(x, num_classes, dtype=_np.float32):
    out = _np.zeros((int(num_classes),), dtype=dtype)
    i = int(x)
    if 0 <= i < int(num_classes):          # jax: out-of-range index -> all zeros
        out[i] = 1
    return out
