"""Shim of the parts of dm-haiku the reference's model uses: ``Module``, ``Linear``, ``transform``,
``without_apply_rng``.  Parameter naming follows haiku (and the shipped ``Test/lunar_lander/params.pickle``):
a module created inside another module's ``__init__`` is named ``<parent>/~/<name>[_k]``."""
import math
from collections import OrderedDict
from typing import Mapping

import numpy as _np

Params = Mapping

_frame = None          # {"params": dict, "init_rng": np Generator or None, "counts": {scope: {base: n}}, "stack": [module names]}


def _unique(scope, base):
    c = _frame["counts"].setdefault(scope, {})
    k = c.get(base, 0)
    c[base] = k + 1
    return base if k == 0 else f"{base}_{k}"


def _snake(name):
    out = ""
    for i, ch in enumerate(name):
        if ch.isupper() and i and not name[i - 1].isupper():
            out += "_"
        out += ch.lower()
    return out


class Module:
    def __init__(self, name=None):
        if _frame is None:
            raise RuntimeError("haiku modules must be created inside hk.transform")
        parent = _frame["constructing"]
        base = name or _snake(type(self).__name__)
        if parent is None:
            self.module_name = _unique("", base)
        else:
            self.module_name = parent + "/~/" + _unique(parent, base)
        if not isinstance(self, Linear):
            _frame["constructing"] = self.module_name      # submodules built in this __init__ nest under it


class Linear(Module):
    def __init__(self, output_size, name=None):
        super().__init__(name)
        self.output_size = int(output_size)

    def __call__(self, x):
        p = _frame["params"]
        if self.module_name not in p:
            if _frame["init_rng"] is None:
                raise KeyError(f"missing parameters for {self.module_name}")
            fan_in = int(x.shape[-1])
            std = 1.0 / math.sqrt(fan_in)
            w = _frame["init_rng"].standard_normal((fan_in, self.output_size))
            w = _np.clip(w, -2.0, 2.0) * std                # TruncatedNormal(stddev = 1/sqrt(fan_in)), +-2 sigma
            p[self.module_name] = {"w": w.astype(_np.float32), "b": _np.zeros(self.output_size, _np.float32)}
        leaves = p[self.module_name]
        return x @ leaves["w"] + leaves["b"]                 # jnp.dot(x, w) + b, w is [in, out]


class Transformed:
    def __init__(self, fun, with_rng=True):
        self._fun, self._with_rng = fun, with_rng

    def _run(self, params, rng, args):
        global _frame
        saved = _frame
        _frame = {"params": params, "init_rng": rng, "counts": {}, "constructing": None}
        try:
            return self._fun(*args)
        finally:
            _frame = saved

    def init(self, rng, *args):
        seed = int(_np.asarray(rng).ravel()[-1]) if not isinstance(rng, (int, _np.integer)) else int(rng)
        params = OrderedDict()
        self._run(params, _np.random.default_rng(seed), args)
        return params

    def apply(self, params, *args):
        if self._with_rng:
            args = args[1:]                                    # apply(params, rng, *args)
        return self._run(params, None, args)


def transform(fun):
    return Transformed(fun, with_rng=True)


def without_apply_rng(t):
    return Transformed(t._fun, with_rng=False)
