"""ctypes binding of libspam_cuda.so (include/spam_cuda.h).  No CPU fallback: if the library or a
CUDA device is missing, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libspam_cuda.so")

SPAM_F32, SPAM_F64, SPAM_I32, SPAM_I64 = 0, 1, 2, 3
DTYPES = {np.dtype(np.float32): SPAM_F32, np.dtype(np.float64): SPAM_F64, np.dtype(np.int32): SPAM_I32,
          np.dtype(np.int64): SPAM_I64}
NP_OF = {v: k for k, v in DTYPES.items()}

STATUS_NAMES = {0: "SPAM_OK", 1: "SPAM_EINVAL", 2: "SPAM_EDIM", 3: "SPAM_ECOLS", 4: "SPAM_ENOMEM", 5: "SPAM_ECUDA",
                6: "SPAM_ESTATE", 7: "SPAM_EOVERFLOW", 8: "SPAM_EINDEX", 9: "SPAM_EDTYPE"}

# every symbol include/spam_cuda.h declares (tests check the .so exports each of them)
EXPORTS = [
    "spam_cuda_create", "spam_cuda_destroy", "spam_cuda_set_stream", "spam_cuda_set_timing", "spam_cuda_get_stats",
    "spam_cuda_synchronize", "spam_cuda_get_phase_totals", "spam_strerror", "spam_last_error", "spam_cuda_abi_version", "spam_host_alloc",
    "spam_host_free", "spam_spgemm_symbolic", "spam_spgemm_numeric", "spam_spmv", "spam_dok_to_csr",
    "spam_dok_to_csr_fetch", "spam_csr_upload", "spam_dcsr_wrap", "spam_dcsr_info", "spam_dcsr_download",
    "spam_dcsr_free", "spam_dcsr_slice_rows", "spam_dcsr_select_rows", "spam_dcsr_transpose", "spam_csr_transpose", "spam_spgemm_dev", "spam_spgemm_dev_b2", "spam_spmv_dev", "spam_dok_to_csr_dev",
    "spam_rows_to_parts", "spam_rows_to_parts_cost", "spam_offset_u64", "spam_dcsr_ewise", "spam_csr_ewise",
    "spam_csr_ewise_fetch", "spam_mm_parse", "spam_mm_free",
    "spam_comm_unique_id", "spam_comm_init", "spam_comm_destroy", "spam_comm_info", "spam_comm_broadcast",
    "spam_comm_allgather_u64", "spam_comm_allgatherv", "spam_spgemm_gathered", "spam_spmv_gathered", "spam_dok_to_csr_sharded",
]


class SpamStats(C.Structure):
    _fields_ = [("flops", C.c_uint64), ("nnz_c", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64), ("ms_flop", C.c_float),
                ("ms_symbolic", C.c_float), ("ms_scan", C.c_float), ("ms_numeric", C.c_float),
                ("ms_total", C.c_float), ("sym_bin_rows", C.c_uint32 * 16), ("num_bin_rows", C.c_uint32 * 16),
                ("fallbacks", C.c_uint32 * 8)]


TINY_BIN, HEAVY_BIN, MERGE_BIN = 0, 9, 10   # indices into sym_bin_rows / num_bin_rows; 1..8 = hash bins


class SpamError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


class DimensionMismatch(SpamError):
    pass


_lib = None


def load():
    """dlopen the in-tree library.  Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: build it with `python -m sparse_matrix_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.spam_strerror.restype = C.c_char_p
    L.spam_strerror.argtypes = [i32]
    L.spam_last_error.restype = C.c_char_p
    L.spam_last_error.argtypes = [vp]
    L.spam_cuda_create.argtypes = [C.POINTER(vp), i32]
    L.spam_cuda_destroy.argtypes = [vp]
    L.spam_cuda_set_stream.argtypes = [vp, vp]
    L.spam_cuda_set_timing.argtypes = [vp, i32]
    L.spam_cuda_get_stats.argtypes = [vp, C.POINTER(SpamStats)]
    L.spam_cuda_synchronize.argtypes = [vp]
    L.spam_cuda_get_phase_totals.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64), i32]
    L.spam_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.spam_host_free.argtypes = [vp]
    L.spam_spgemm_symbolic.argtypes = [vp, i32, u64, u64, vp, vp, vp, u64, u64, vp, vp, vp, vp, C.POINTER(u64)]
    L.spam_spgemm_numeric.argtypes = [vp, vp, vp, i32]
    L.spam_spmv.argtypes = [vp, i32, u64, u64, vp, vp, vp, vp, vp]
    L.spam_dok_to_csr.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp, vp, C.POINTER(u64)]
    L.spam_dok_to_csr_fetch.argtypes = [vp, vp, vp]
    L.spam_csr_upload.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp, C.POINTER(vp)]
    L.spam_dcsr_wrap.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp, C.POINTER(vp)]
    L.spam_dcsr_info.argtypes = [vp, C.POINTER(i32), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(vp),
                                 C.POINTER(vp), C.POINTER(vp)]
    L.spam_dcsr_download.argtypes = [vp, vp, vp, vp, vp]
    L.spam_dcsr_free.argtypes = [vp, vp]
    L.spam_dcsr_slice_rows.argtypes = [vp, vp, u64, u64, C.POINTER(vp)]
    L.spam_dcsr_select_rows.argtypes = [vp, vp, vp, u64, C.POINTER(vp)]
    L.spam_dcsr_ewise.argtypes = [vp, i32, vp, vp, C.POINTER(vp)]
    L.spam_csr_ewise.argtypes = [vp, i32, i32, u64, u64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(u64)]
    L.spam_csr_ewise_fetch.argtypes = [vp, vp, vp]
    L.spam_dcsr_transpose.argtypes = [vp, vp, C.POINTER(vp)]
    L.spam_csr_transpose.argtypes = [vp, i32, u64, u64, vp, vp, vp, vp, vp, vp]
    L.spam_spgemm_dev.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.spam_spgemm_dev_b2.argtypes = [vp, vp, vp, i32, C.POINTER(vp)]
    L.spam_spmv_dev.argtypes = [vp, vp, vp, vp]
    L.spam_dok_to_csr_dev.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp, C.POINTER(vp)]
    L.spam_rows_to_parts.argtypes = [vp, vp, vp, C.c_uint32, vp, C.POINTER(u64)]
    L.spam_rows_to_parts_cost.argtypes = [vp, vp, vp, C.c_uint32, vp, C.POINTER(u64)]
    L.spam_offset_u64.argtypes = [vp, vp, u64, u64]
    L.spam_comm_unique_id.argtypes = [vp]
    L.spam_comm_init.argtypes = [vp, vp, i32, i32]
    L.spam_comm_destroy.argtypes = [vp]
    L.spam_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.spam_comm_broadcast.argtypes = [vp, vp, u64, i32]
    L.spam_comm_allgather_u64.argtypes = [vp, vp, C.c_uint32, vp]
    L.spam_comm_allgatherv.argtypes = [vp, vp, vp]
    L.spam_spgemm_gathered.argtypes = [vp, vp, vp, u64, u64, i32, i32, C.POINTER(vp)]
    L.spam_spmv_gathered.argtypes = [vp, vp, vp, vp, vp]
    L.spam_dok_to_csr_sharded.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp, C.POINTER(u64), C.POINTER(vp)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("spam_strerror", "spam_last_error", "spam_mm_free"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(handle_ptr, status: int):
    if status == 0:
        return
    L = load()
    msg = L.spam_strerror(status).decode()
    if handle_ptr:
        detail = L.spam_last_error(handle_ptr).decode()
        if detail:
            msg = f"{msg} ({detail})"
    if status == 2:
        raise DimensionMismatch(status, msg)
    if status == 8:
        raise IndexError(f"IndexError: {msg}")
    raise SpamError(status, msg)


def ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(None)
    return C.c_void_p(a.ctypes.data)
