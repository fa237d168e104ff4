import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def built():
    """The product library, kernel images and CPU checkers are built in-tree by __graft_entry__.build();
    build them if a fresh checkout has none (never silently fall back to anything else)."""
    lib = os.path.join(ROOT, "shaderbox_b200", "libsbx.so")
    util = os.path.join(ROOT, "shaderbox_b200", "images", "sbx_util.cubin")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(util) and os.path.exists(orc)):
        import __graft_entry__

        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def golden_frames():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))


@pytest.fixture(scope="session")
def golden_ops():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))
