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


@pytest.fixture(scope="session")
def handle_nosort():
    """A handle created with SPAM_SORT_B=0: an unsorted right-hand side is multiplied as it is (thread-per-row and
    warp hash bins) instead of through its cached sorted copy — the tests that pin those bins use it."""
    import sparse_matrix_b200 as S
    os.environ["SPAM_SORT_B"] = "0"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_SORT_B"]
    yield h
    h.close()


@pytest.fixture(scope="session")
def handle_nobucket():
    """SPAM_DOK_BUCKET=0: DOK -> CSR and transpose by the counting / radix paths of dok.cu, which the bucket path
    (bucket.cuh, the default) falls back to for shapes it does not take."""
    import sparse_matrix_b200 as S
    os.environ["SPAM_DOK_BUCKET"] = "0"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_DOK_BUCKET"]
    yield h
    h.close()


@pytest.fixture(scope="session")
def handle_msort():
    """SPAM_ESC=3: rows that do not compress take the merge-tree bins (msort.cuh) instead of the hash bins."""
    import sparse_matrix_b200 as S
    os.environ["SPAM_ESC"] = "3"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_ESC"]
    yield h
    h.close()


@pytest.fixture(scope="session")
def handle_esc():
    """SPAM_ESC=2 (and SPAM_SORT_B=0): rows that do not compress take the bucket-sort bins 11..15 (esc.cuh), which are
    off by default because the hash bins measured faster on B200."""
    import sparse_matrix_b200 as S
    os.environ["SPAM_SORT_B"] = "0"
    os.environ["SPAM_ESC"] = "2"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_SORT_B"]
        del os.environ["SPAM_ESC"]
    yield h
    h.close()


@pytest.fixture(scope="session")
def handle_win():
    """SPAM_MERGE_WIN=3: the merge-bin kernels that stage the block's window of B in shared memory with cp.async.bulk
    (k_flop_sym_merge_win / k_num_merge_win, merge.cuh) — off by default because they measured slower on B200."""
    import sparse_matrix_b200 as S
    os.environ["SPAM_MERGE_WIN"] = "3"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_MERGE_WIN"]
    yield h
    h.close()
