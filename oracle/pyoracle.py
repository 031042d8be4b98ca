"""ctypes loader for the CPU oracle (oracle/spam_oracle.cpp).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (sparse_matrix_b200) never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libspam_oracle.so")

DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
      np.dtype(np.int8): 4}

_u64p = C.POINTER(C.c_uint64)
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "spam_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libspam_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_mul_hash.restype = C.c_int
        L.oracle_rows_to_threads.restype = C.c_int
        L.oracle_symbolic.restype = C.c_int
        L.oracle_dok_to_csr.restype = C.c_int
        L.oracle_dok_dense_mul.restype = C.c_int
        L.oracle_spmv.restype = C.c_int
        L.oracle_transpose.restype = C.c_int
        L.oracle_ewise.restype = C.c_int
        L.oracle_table_size_for.restype = C.c_uint64
        L.oracle_table_size_for.argtypes = [C.c_uint64]
        L.oracle_hash.restype = C.c_uint64
        L.oracle_hash.argtypes = [C.c_uint32]
        L.oracle_hashset_run.restype = C.c_uint64
        L.oracle_hashset_run2.restype = C.c_uint64
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_hardware_threads.restype = C.c_uint
        _lib = L
    return _lib


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _take(ptr, n, dtype):
    """copy n items out of a malloc'd buffer and free it"""
    L = lib()
    if n:
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        out = np.frombuffer(buf, dtype=dtype, count=n).copy()
    else:
        out = np.empty(0, dtype=dtype)
    L.oracle_free(C.c_void_p(ptr))
    return out


def mul_hash(a, b, sorted_output: bool, tnum: int = 0, want_flops: bool = False):
    """a, b: (rows, cols, offsets, indices, vals) tuples.  Returns (offsets, indices, vals[, flops])
    exactly as spam_csr::CsrMatrix::mul_hash::<_, B2=sorted_output> would."""
    L = lib()
    ar, ac, ao, ai, av = a
    br, bc, bo, bi, bv = b
    av = np.ascontiguousarray(av)
    bv = np.ascontiguousarray(bv, dtype=av.dtype)
    ao, ai, bo, bi = _u64(ao), _u64(ai), _u64(bo), _u64(bi)
    dt = DT[av.dtype]
    c_off, c_idx, c_val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz, flops = C.c_uint64(), C.c_uint64()
    rc = L.oracle_mul_hash(C.c_int(dt), C.c_uint64(ar), C.c_uint64(ac), _ptr(ao), _ptr(ai), _ptr(av),
                           C.c_uint64(br), C.c_uint64(bc), _ptr(bo), _ptr(bi), _ptr(bv), C.c_int(int(sorted_output)),
                           C.c_uint64(tnum), C.byref(c_off), C.byref(c_idx), C.byref(c_val), C.byref(nnz),
                           C.byref(flops))
    if rc != 0:
        raise RuntimeError(f"oracle_mul_hash failed rc={rc}")
    n = nnz.value
    off = _take(c_off.value, ar + 1, np.uint64)
    idx = _take(c_idx.value, n, np.uint64)
    val = _take(c_val.value, n, av.dtype)
    if want_flops:
        return off, idx, val, flops.value
    return off, idx, val


def mul_hash_timed(a, b, sorted_output: bool, tnum: int = 0):
    """Runs the oracle product, frees the result, returns (seconds, nnz, flops) — for the CPU baseline."""
    import time
    L = lib()
    ar, ac, ao, ai, av = a
    br, bc, bo, bi, bv = b
    dt = DT[av.dtype]
    c_off, c_idx, c_val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz, flops = C.c_uint64(), C.c_uint64()
    t0 = time.perf_counter()
    rc = L.oracle_mul_hash(C.c_int(dt), C.c_uint64(ar), C.c_uint64(ac), _ptr(ao), _ptr(ai), _ptr(av),
                           C.c_uint64(br), C.c_uint64(bc), _ptr(bo), _ptr(bi), _ptr(bv), C.c_int(int(sorted_output)),
                           C.c_uint64(tnum), C.byref(c_off), C.byref(c_idx), C.byref(c_val), C.byref(nnz),
                           C.byref(flops))
    t1 = time.perf_counter()
    if rc != 0:
        raise RuntimeError(f"oracle_mul_hash failed rc={rc}")
    for p in (c_off, c_idx, c_val):
        L.oracle_free(p)
    return t1 - t0, nnz.value, flops.value


def rows_to_threads(a_rows, a_off, a_idx, b_off, tnum):
    L = lib()
    a_off, a_idx, b_off = _u64(a_off), _u64(a_idx), _u64(b_off)
    flop = np.zeros(a_rows, dtype=np.uint64)
    ro = np.zeros(tnum + 1, dtype=np.uint64)
    rc = L.oracle_rows_to_threads(C.c_uint64(a_rows), _ptr(a_off), _ptr(a_idx), _ptr(b_off), C.c_uint64(tnum),
                                  _ptr(flop), _ptr(ro))
    if rc != 0:
        raise RuntimeError(f"oracle_rows_to_threads rc={rc}")
    return flop, ro


def symbolic(a_rows, a_off, a_idx, b_off, b_idx, tnum=0):
    L = lib()
    a_off, a_idx, b_off, b_idx = _u64(a_off), _u64(a_idx), _u64(b_off), _u64(b_idx)
    out = np.zeros(a_rows, dtype=np.uint64)
    rc = L.oracle_symbolic(C.c_uint64(a_rows), _ptr(a_off), _ptr(a_idx), _ptr(b_off), _ptr(b_idx), C.c_uint64(tnum),
                           _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle_symbolic rc={rc}")
    return out


def dok_to_csr(rows, cols, ri, ci, vals):
    """Sequential DokMatrix::set_element over the triplet stream, then CsrMatrix::from(dok)."""
    L = lib()
    vals = np.ascontiguousarray(vals)
    ri, ci = _u64(ri), _u64(ci)
    c_off, c_idx, c_val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz = C.c_uint64()
    rc = L.oracle_dok_to_csr(C.c_int(DT[vals.dtype]), C.c_uint64(rows), C.c_uint64(cols), C.c_uint64(len(vals)),
                             _ptr(ri), _ptr(ci), _ptr(vals), C.byref(c_off), C.byref(c_idx), C.byref(c_val),
                             C.byref(nnz))
    if rc == 4:
        raise IndexError("IndexError")
    if rc != 0:
        raise RuntimeError(f"oracle_dok_to_csr rc={rc}")
    n = nnz.value
    return _take(c_off.value, rows + 1, np.uint64), _take(c_idx.value, n, np.uint64), _take(c_val.value, n, vals.dtype)


def transpose(mat, literal: bool = False):
    """CsrMatrix::transpose (lib.rs:256-264): (offsets, indices, vals) of the cols x rows result.  `literal`
    runs the reference's own (j, i) double loop (small inputs only)."""
    L = lib()
    rows, cols, off, idx, val = mat
    off, idx, val = _u64(off), _u64(idx), np.ascontiguousarray(val)
    nnz = int(off[rows])
    t_off = np.zeros(cols + 1, np.uint64)
    t_idx = np.zeros(nnz, np.uint64)
    t_val = np.zeros(nnz, val.dtype)
    rc = L.oracle_transpose(C.c_int(DT[val.dtype]), C.c_uint64(rows), C.c_uint64(cols), _ptr(off), _ptr(idx), _ptr(val),
                            C.c_int(1 if literal else 0), _ptr(t_off), _ptr(t_idx), _ptr(t_val))
    if rc == 4:
        raise IndexError("IndexError")
    if rc != 0:
        raise RuntimeError(f"oracle_transpose rc={rc}")
    return t_off, t_idx, t_val


def ewise(a, b, op: str = "add", is_sorted: bool = True):
    """impl Add / Sub for CsrMatrix (apply_elementwise, lib.rs:83-149): (offsets, indices, vals) of a op b.
    `is_sorted` is the operands' IS_SORTED const generic (selects the merge-join or the HashMap branch)."""
    L = lib()
    ar, ac, a_off, a_idx, a_val = a
    br, bc, b_off, b_idx, b_val = b
    a_off, a_idx, a_val = _u64(a_off), _u64(a_idx), np.ascontiguousarray(a_val)
    b_off, b_idx, b_val = _u64(b_off), _u64(b_idx), np.ascontiguousarray(b_val, dtype=a_val.dtype)
    c_off, c_idx, c_val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz = C.c_uint64()
    rc = L.oracle_ewise(C.c_int(DT[a_val.dtype]), C.c_int({"add": 0, "sub": 1}[op]), C.c_int(1 if is_sorted else 0),
                        C.c_uint64(ar), C.c_uint64(ac), C.c_uint64(br), C.c_uint64(bc), _ptr(a_off), _ptr(a_idx),
                        _ptr(a_val), _ptr(b_off), _ptr(b_idx), _ptr(b_val), C.byref(c_off), C.byref(c_idx),
                        C.byref(c_val), C.byref(nnz))
    if rc == 2:
        raise ValueError("matrices must have identical dimensions")
    if rc != 0:
        raise RuntimeError(f"oracle_ewise rc={rc}")
    n = nnz.value
    return _take(c_off.value, ar + 1, np.uint64), _take(c_idx.value, n, np.uint64), _take(c_val.value, n, a_val.dtype)


def dok_dense_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Dense triple loop of `impl Mul for &DokMatrix` (zeros = absent entries)."""
    L = lib()
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    l, m = a.shape
    m2, n = b.shape
    assert m == m2, "LHS cols != RHS rows"
    c = np.zeros((l, n), dtype=a.dtype)
    rc = L.oracle_dok_dense_mul(C.c_int(DT[a.dtype]), C.c_uint64(l), C.c_uint64(m), C.c_uint64(n), _ptr(a), _ptr(b),
                                _ptr(c))
    assert rc == 0
    return c


def spmv(rows, cols, off, idx, val, x):
    L = lib()
    val = np.ascontiguousarray(val)
    x = np.ascontiguousarray(x, dtype=val.dtype)
    off, idx = _u64(off), _u64(idx)
    y = np.zeros(rows, dtype=val.dtype)
    rc = L.oracle_spmv(C.c_int(DT[val.dtype]), C.c_uint64(rows), C.c_uint64(cols), _ptr(off), _ptr(idx), _ptr(val),
                       _ptr(x), _ptr(y))
    assert rc == 0
    return y


def table_size_for(cap: int) -> int:
    return lib().oracle_table_size_for(cap)


def hash_of(key: int) -> int:
    return lib().oracle_hash(key)


def hashset_run(keys, cap_hint=0, slots_cap=4096):
    L = lib()
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    ub = C.c_uint64()
    slots = np.zeros(slots_cap, dtype=np.uint32)
    n = L.oracle_hashset_run(_ptr(keys), C.c_uint64(len(keys)), C.c_uint64(cap_hint), C.byref(ub), _ptr(slots),
                             C.c_uint64(slots_cap))
    return n, ub.value, slots[: min(slots_cap, ub.value)].copy()


def hashset_run2(keys, initial_capacity, shrink_cap=0, slots_cap=4096):
    """HashSet::with_capacity(initial_capacity), optional shrink_to(shrink_cap), then the inserts.
    Returns (len, upper_bound, allocated slots, slots[:upper_bound])."""
    L = lib()
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    ub, alloc = C.c_uint64(), C.c_uint64()
    slots = np.zeros(slots_cap, dtype=np.uint32)
    n = L.oracle_hashset_run2(_ptr(keys), C.c_uint64(len(keys)), C.c_uint64(initial_capacity), C.c_uint64(shrink_cap),
                              C.byref(ub), C.byref(alloc), _ptr(slots), C.c_uint64(slots_cap))
    return n, ub.value, alloc.value, slots[: min(slots_cap, ub.value)].copy()


def hardware_threads() -> int:
    return lib().oracle_hardware_threads()
