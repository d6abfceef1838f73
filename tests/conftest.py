import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container (GPU tests run on the B200 box)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def syn():
    import mc3d_b200.synthetic as s
    return s


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def cams_from_golden(g, n):
    return {i: [g[f'cam{i}_K'], g[f'cam{i}_R'], g[f'cam{i}_T'], g[f'cam{i}_dist']] for i in range(n)}


def rel_err(a, b):
    """Per-point ||a-b|| / ||b|| over the last axis."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, a.shape[-1])
    b = np.asarray(b, dtype=np.float64).reshape(-1, b.shape[-1])
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
