import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def models():
    return np.load(os.path.join(GOLDEN, "ref_models.npz"))


@pytest.fixture(scope="session")
def ref_dynamics():
    return np.load(os.path.join(GOLDEN, "ref_dynamics.npz"))


@pytest.fixture(scope="session")
def costmap():
    from autorally_b200.params import make_ellipse_costmap
    return make_ellipse_costmap()


@pytest.fixture(scope="session")
def small_costmap():
    """Coarser ellipse map (5 px/m) for the fast CPU tests."""
    from autorally_b200.params import make_ellipse_costmap
    return make_ellipse_costmap(pixels_per_meter=5.0)
