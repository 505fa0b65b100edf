import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "golden.npz")
CALIB = os.path.join(ROOT, "tests", "golden", "zed2_calibration.yaml")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: a few seconds of scalar CPU work")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def calib_path():
    return CALIB


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU session. Fails (does not skip) when the library or device is missing."""
    from stereo_depth_ruler_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.build()
    return o
