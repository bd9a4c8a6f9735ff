import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore", message=".*torch.meshgrid.*")

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (nvcc cross-compiles without a GPU)."""
    from multiviewhmr_b200 import build
    return build.build()


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "models"))


@pytest.fixture(scope="session")
def reference():
    """The real reference modules, only where /root/reference is mounted."""
    if not have_reference():
        pytest.skip("/root/reference not present on this machine")
    import importlib
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("models", "utils") or k.startswith(("models.", "utils."))}
    sys.path.insert(0, REFERENCE)
    try:
        agg = importlib.import_module("models.aggregation")
        mv = importlib.import_module("utils.multiview")
        vol = importlib.import_module("utils.volumetric")
    finally:
        sys.path.remove(REFERENCE)
    mods = {"aggregation": agg, "multiview": mv, "volumetric": vol}
    for k in list(sys.modules):
        if k in ("models", "utils") or k.startswith(("models.", "utils.")):
            sys.modules.pop(k)
    sys.modules.update(saved)
    return mods


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
