"""sparse_matrix_b200 — B200-native (sm_100a) SpGEMM / SpMV / DOK->CSR behind the reference's
spam_csr operator interface.  The compute lives in libspam_cuda.so (hand-written CUDA, C ABI in
include/spam_cuda.h); this package is the host-side mirror of the reference API over that ABI."""
from ._lib import SO_PATH, DimensionMismatch, SpamError, load  # noqa: F401
from .csr import CsrMatrix, DeviceCsr, DokMatrix, Handle, comm_unique_id, get_handle  # noqa: F401
from .matrix_market import (FromMatrixMarketError, into_float_matrix_market, load_matrix_market,  # noqa: F401
                            parse_matrix_market)

__all__ = ["CsrMatrix", "DokMatrix", "DeviceCsr", "Handle", "get_handle", "SpamError", "DimensionMismatch", "load",
           "SO_PATH"]
