import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_seed1.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def init_cells(golden):
    """fresh reference world, SEED=1: tiled AoS cell pool with the heights of map::init"""
    import numpy as np
    import orc
    cells = np.zeros(512 * 512, orc.CELL_DTYPE)
    cells["height"] = golden["init_height"]
    return cells
