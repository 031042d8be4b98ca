import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/spam_oracle.cpp) — the checker, never the thing under test in -m gpu tests."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def handle():
    """spam_handle on cuda:0.  Fails loudly when the CUDA library or device is missing."""
    import sparse_matrix_b200 as S
    h = S.get_handle(0)
    yield h
