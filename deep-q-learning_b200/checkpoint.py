"""Checkpoint I/O in the reference's on-disk layout (``General/Base/utils.py:21-40``).

``generate_saving(directory)`` / ``generate_loading(directory)`` keep the reference's names and file
names (``params.pickle``, ``opt_state.pickle``).  Trees are written as plain numpy arrays inside the
same nesting (``{module: {'w','b'}}`` and ``(ScaleByAdamState(count, mu, nu), EmptyState(), ...)``);
the loader additionally understands the reference's own pickles, whose leaves are
``jax._src.device_array.reconstruct_device_array`` records and whose optimiser state classes live in
``optax._src`` (neither library is importable here) -- e.g. ``Test/lunar_lander/*.pickle``.
"""
import os
import pickle
from collections import OrderedDict

import numpy as np

from .specs import EmptyState, ScaleByAdamState


def _reconstruct_device_array(fun, args, arr_state, aval_state):
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return arr


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("jax") and name == "reconstruct_device_array":
            return _reconstruct_device_array
        if module.startswith("optax") and name == "ScaleByAdamState":
            return ScaleByAdamState
        if module.startswith("optax") and name == "EmptyState":
            return EmptyState
        return super().find_class(module, name)


def _to_numpy_tree(tree):
    return OrderedDict((m, {k: np.asarray(v) for k, v in leaves.items()}) for m, leaves in tree.items())


def load_pickle(path):
    with open(path, "rb") as f:
        return _Unpickler(f).load()


def generate_saving(directory):
    def save_state(params, opt_state):
        if not os.path.exists(directory):
            os.mkdir(directory)
        with open(os.path.join(directory, "params.pickle"), "wb") as f:
            pickle.dump(_to_numpy_tree(params), f)
        first = opt_state[0]
        adam_state = ScaleByAdamState(np.asarray(first[0], dtype=np.int32), _to_numpy_tree(first[1]),
                                      _to_numpy_tree(first[2]))
        with open(os.path.join(directory, "opt_state.pickle"), "wb") as f:
            pickle.dump((adam_state,) + tuple(opt_state[1:]), f)
    return save_state


def generate_loading(directory):
    def load_state():
        params = _to_numpy_tree(load_pickle(os.path.join(directory, "params.pickle")))
        opt_state = load_pickle(os.path.join(directory, "opt_state.pickle"))
        first = opt_state[0]
        adam_state = ScaleByAdamState(np.asarray(first[0], dtype=np.int32), _to_numpy_tree(first[1]),
                                      _to_numpy_tree(first[2]))
        return params, (adam_state,) + tuple(opt_state[1:])
    return load_state
