import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun on a B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    class G:
        positions = np.load(os.path.join(GOLDEN, "positions.npz"))
        playouts = np.load(os.path.join(GOLDEN, "playouts.npz"))
        mcts = np.load(os.path.join(GOLDEN, "mcts.npz"))
        import json
        kats = json.load(open(os.path.join(GOLDEN, "kats.json"), encoding="utf-8"))
        end_reasons = json.load(open(os.path.join(GOLDEN, "end_reasons.json"), encoding="utf-8"))
    return G


@pytest.fixture(scope="session")
def xo():
    from oracle import xq_oracle
    xq_oracle.build()
    return xq_oracle


@pytest.fixture(scope="session")
def built_lib():
    """libxq_b200.so, compiling it with nvcc if this checkout has not built it yet."""
    from chinesechessai_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()
