import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/ebur128_oracle.c) behind the shared ctypes binding."""
    from oracle import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def product():
    """The product library (CUDA).  Only meaningful in -m gpu tests."""
    from loudgain_b200 import build, load_library
    build()
    return load_library()


@pytest.fixture(scope="session")
def product_path():
    """Path of the built product library (building needs no GPU)."""
    from loudgain_b200 import build
    return build()
