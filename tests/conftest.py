import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fx():
    """Golden fixtures extracted from the reference's data files (tests/golden/make_fixtures.py)."""
    path = os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")
    return dict(np.load(path, allow_pickle=False))
