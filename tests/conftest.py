import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected explicitly with ``-m gpu``; on a box without a device they are
    # skipped (the product itself never falls back -- it raises).
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


@pytest.fixture(scope="session")
def ref_windows():
    return np.load(os.path.join(GOLDEN, "ref_windows.npz"))


@pytest.fixture(scope="session")
def ref_spectral():
    return np.load(os.path.join(GOLDEN, "ref_spectral.npz"))


@pytest.fixture(scope="session")
def ref_location():
    return np.load(os.path.join(GOLDEN, "ref_location.npz"))


def split_feature(key):
    """'percentile:90' -> ('percentile', 90.0); 'mean' -> ('mean', None)."""
    if ":" in key:
        n, p = key.split(":")
        return n, float(p)
    return key, None


WINDOW_CASES = ["acc_z_500_250", "acc_x_500_250", "ppg_1920_64", "acc_y_64_48", "acc_z_7_3",
                "quant_50_10", "exact_fit_100_100", "single_window"]
