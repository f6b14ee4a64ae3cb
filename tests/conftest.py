import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ert():
    import eraytracer_b200
    eraytracer_b200.load()          # raises if the shared object is missing: no fallback
    return eraytracer_b200


@pytest.fixture(scope="session")
def gpu(ert):
    n = ert.device_count()          # raises ErtError when no device works
    assert n >= 1
    return n
