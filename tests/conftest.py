import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), built on demand."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_scoring():
    return dict(np.load(os.path.join(GOLDEN, "golden_scoring.npz")))


@pytest.fixture(scope="session")
def golden_primitives():
    return dict(np.load(os.path.join(GOLDEN, "golden_primitives.npz")))


@pytest.fixture(scope="session")
def golden_scoring_sphere():
    return dict(np.load(os.path.join(GOLDEN, "golden_scoring_sphere.npz")))


@pytest.fixture(scope="session")
def golden_primitives_wide():
    return dict(np.load(os.path.join(GOLDEN, "golden_primitives_wide.npz")))


@pytest.fixture(scope="session")
def golden_views(orc, golden_scoring):
    g = golden_scoring
    return orc.Views(g["P"], list(g["images"]))
