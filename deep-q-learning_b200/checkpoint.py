"""Checkpoint I/O in the reference's on-disk layout (``General/Base/utils.py:21-40``).

``generate_saving(directory)`` / ``generate_loading(directory)`` keep the reference's names and file
names (``params.pickle``, ``opt_state.pickle``).  Trees are written as plain numpy arrays inside the
same nesting (``{module: {'w','b'}}`` and ``(ScaleByAdamState(count, mu, nu), EmptyState(), ...)``).
The two optimiser-state classes are written under the names the REFERENCE's pickles carry
(``optax._src.transform.ScaleByAdamState``, ``optax._src.base.EmptyState``), not under this
package's module path, so the reference's ``generate_loading`` -- a bare ``pickle.load`` in a
process that has optax but has never heard of this package -- reads the files
(``tests/test_host_logic.py::test_checkpoint_loads_with_bare_pickle``).  The loader additionally understands the reference's own pickles, whose leaves are
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


# class -> (module, qualified name) the reference's pickles use for it (Test/lunar_lander/opt_state.pickle)
_REFERENCE_GLOBALS = {ScaleByAdamState: ("optax._src.transform", "ScaleByAdamState"),
                      EmptyState: ("optax._src.base", "EmptyState")}


class _ReferencePickler(pickle._Pickler):
    """Pure-Python pickler (protocol 4, like the fixture) that names the optimiser-state classes the way
    optax does.  ``save_global`` normally insists that the named module is importable HERE; optax is not,
    so the GLOBAL record is written directly."""

    def save_global(self, obj, name=None):
        target = _REFERENCE_GLOBALS.get(obj)
        if target is None:
            return super().save_global(obj, name)
        self.save(target[0])
        self.save(target[1])
        self.write(pickle.STACK_GLOBAL)
        self.memoize(obj)


def _dump(obj, f):
    _ReferencePickler(f, protocol=4).dump(obj)


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
            _dump(dict(_to_numpy_tree(params)), f)
        first = opt_state[0]
        adam_state = ScaleByAdamState(np.asarray(first[0], dtype=np.int32), dict(_to_numpy_tree(first[1])),
                                      dict(_to_numpy_tree(first[2])))
        rest = tuple(EmptyState() if isinstance(x, tuple) and len(x) == 0 else x for x in opt_state[1:])
        with open(os.path.join(directory, "opt_state.pickle"), "wb") as f:
            _dump((adam_state,) + rest, f)
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
