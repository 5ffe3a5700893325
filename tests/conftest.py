import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_deviations_test():
    return _load("deviations_test.npz")


@pytest.fixture(scope="session")
def golden_mini():
    return _load("mini.npz")


@pytest.fixture(scope="session")
def golden_example_counts():
    return _load("example_counts.npz")


@pytest.fixture(scope="session")
def engine():
    """The process-wide cv_handle on cuda:0.  Fails loudly when the CUDA library or GPU is missing."""
    from chromvar_b200 import api
    return api.get_engine()
