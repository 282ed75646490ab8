import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer CPU test")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


C1_SCATTERERS = [(8.0, 0.0, -10.0), (12.0, 30.0, -8.0), (16.0, -20.0, -6.0), (20.0, 10.0, -3.0), (25.0, -40.0, 0.0)]


def c1_scatterers():
    """BASELINE.md section 3 scene: rows (range m, azimuth rad, rcs dB, vr)."""
    import numpy as np
    return np.array([(r, np.radians(a), s, 0.0) for r, a, s in C1_SCATTERERS])
