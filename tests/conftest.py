import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference path (test infrastructure only)."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    import orc_b200
    return orc_b200.default_context(0)


def rel_l2(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a - b)


def max_rel(a, b):
    """max |a-b| / max|b| : the 'relative' of the 1e-12 coefficient bar."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.abs(b).max() if b.size else 0.0
    e = np.abs(a - b).max() if b.size else 0.0
    return e / s if s > 0 else e
