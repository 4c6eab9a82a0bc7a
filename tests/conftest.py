import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return {name[:-4]: np.load(os.path.join(GOLDEN, name), allow_pickle=False)
            for name in os.listdir(GOLDEN) if name.endswith(".npz")}


def golden_tree(ck, prefix):
    """Rebuild a {module: {'w','b'}} tree from tests/golden/ref_checkpoint.npz."""
    from collections import OrderedDict
    tree = OrderedDict()
    for mod in [str(m) for m in ck["module_order"]]:
        tree[mod] = {k: ck[f"{prefix}|{mod}|{k}"] for k in ("w", "b")}
    return tree


def assert_close(got, ref, rtol=1e-5, atol_scale=1e-6, what=""):
    """|got - ref| <= rtol*|ref| + atol_scale*max|ref|  -- the north_star's 1e-5 relative fp32 bar, with
    a tiny absolute floor (1e-6 of the tensor's largest magnitude) for entries that are ~0 by cancellation."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    tol = rtol * np.abs(ref) + atol_scale * (np.max(np.abs(ref)) if ref.size else 0.0)
    err = np.abs(got - ref)
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{ref.size} entries off; worst abs err {err.max():.3e} "
                           f"at ref={ref.flat[int(np.argmax(err - tol))]:.6e}")
