"""Import shim: ``import dqn_b200`` loads the package directory ``deep-q-learning_b200/`` (whose name
is not a valid Python identifier) and registers it as ``deep_q_learning_b200``."""
import importlib.util
import os
import sys

_NAME = "deep_q_learning_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deep-q-learning_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
globals().update({k: getattr(pkg, k) for k in pkg.__all__})
__all__ = list(pkg.__all__) + ["pkg"]
