import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def c2a():
    from c2a_loader import c2a as pkg
    return pkg


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def ctx(c2a):
    """Device context. No skip: without a GPU this raises, which is the intended loud failure."""
    return c2a.DeviceContext(0)
