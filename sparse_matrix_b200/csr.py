"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI.

Mirrors (names, argument meaning, error behaviour):
  CsrMatrix<T, IS_SORTED>          spam_csr/src/lib.rs:25-32   fields rows, cols, vals, indices, offsets
  Matrix::invariants               spam_csr/src/lib.rs:47-81,152-160
  CsrMatrix::mul_hash<B1,B2>       spam_csr/src/mul_hash.rs:13-36
  impl Mul for &CsrMatrix          spam_csr/src/lib.rs:292-297  (`a * b`, Output = CsrMatrix<T,false>)
  impl From<DokMatrix> for Csr     spam_csr/src/lib.rs:315-334  (CsrMatrix.from_dok / from_triplets)
  DokMatrix::set_element           spam_dok/src/lib.rs:167-176  (zero removes, else insert/replace)
  IndexError                       spam_matrix/src/lib.rs:13
  spmv                             NEW (no SpMV in the reference; semantics of mul_hash with an n x 1 rhs)

All arithmetic runs in libspam_cuda.so on the GPU.  There is no CPU implementation of the hot path
in this package: without the built library or without a CUDA device every product raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import DTYPES, NP_OF, SpamStats, check, ptr

_handles = {}


class Handle:
    """One stream + workspace on one GPU (spam_handle).  Not thread-safe."""

    def __init__(self, device: int = 0):
        self.L = _lib.load()
        h = C.c_void_p()
        st = self.L.spam_cuda_create(C.byref(h), device)
        if st != 0:
            raise _lib.SpamError(st, f"spam_cuda_create(device={device}) failed: {self.L.spam_strerror(st).decode()} "
                                     "(a CUDA device is required; there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.L.spam_cuda_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        check(self.h, self.L.spam_cuda_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_timing(self, on: bool):
        check(self.h, self.L.spam_cuda_set_timing(self.h, int(on)))

    def synchronize(self):
        check(self.h, self.L.spam_cuda_synchronize(self.h))

    def stats(self) -> dict:
        s = SpamStats()
        check(self.h, self.L.spam_cuda_get_stats(self.h, C.byref(s)))
        return {"flops": s.flops, "nnz_c": s.nnz_c, "kernel_launches": s.kernel_launches, "bytes_h2d": s.bytes_h2d,
                "bytes_d2h": s.bytes_d2h, "ms_flop": s.ms_flop, "ms_symbolic": s.ms_symbolic, "ms_scan": s.ms_scan,
                "ms_numeric": s.ms_numeric, "ms_total": s.ms_total, "sym_bin_rows": list(s.sym_bin_rows),
                "num_bin_rows": list(s.num_bin_rows), "fallbacks": list(s.fallbacks)}


def _comm_unique_id() -> bytes:
    """128-byte rendezvous id (ncclGetUniqueId): rank 0 makes it, every rank passes it to Handle.comm_init."""
    buf = C.create_string_buffer(128)
    st = _lib.load().spam_comm_unique_id(buf)
    if st != 0:
        raise _lib.SpamError(st, "spam_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


def _comm_init(self, unique_id: bytes, rank: int, world: int):
    """Collective: one NCCL communicator inside the library for this handle (spam_comm_init)."""
    assert len(unique_id) == 128
    check(self.h, self.L.spam_comm_init(self.h, C.c_char_p(unique_id), rank, world))
    self.rank, self.world = rank, world


def _comm_info(self) -> dict:
    r, w, p = C.c_int(), C.c_int(), C.c_int()
    check(self.h, self.L.spam_comm_info(self.h, C.byref(r), C.byref(w), C.byref(p)))
    return {"rank": r.value, "world": w.value, "peer_mapped": bool(p.value)}


def _comm_broadcast(self, d_ptr: int, nbytes: int, root: int = 0):
    check(self.h, self.L.spam_comm_broadcast(self.h, C.c_void_p(d_ptr), nbytes, root))


def _comm_allgather_u64(self, values) -> np.ndarray:
    mine = np.ascontiguousarray(values, dtype=np.uint64)
    out = np.zeros(self.world * mine.shape[0], dtype=np.uint64)
    check(self.h, self.L.spam_comm_allgather_u64(self.h, ptr(mine), mine.shape[0], ptr(out)))
    return out.reshape(self.world, mine.shape[0])


def _phase_totals(self, reset: bool = False) -> dict:
    """Sums of the per-phase CUDA-event times (ms) over the products finished since the last reset."""
    ms = (C.c_double * 5)()
    n = C.c_uint64()
    check(self.h, self.L.spam_cuda_get_phase_totals(self.h, ms, C.byref(n), 1 if reset else 0))
    return {"ms_flop": ms[0], "ms_symbolic": ms[1], "ms_scan": ms[2], "ms_numeric": ms[3], "ms_total": ms[4],
            "products": n.value}


Handle.phase_totals = _phase_totals
Handle.comm_init = _comm_init
Handle.comm_info = _comm_info
Handle.comm_broadcast = _comm_broadcast
Handle.comm_allgather_u64 = _comm_allgather_u64
comm_unique_id = _comm_unique_id


def get_handle(device: int = 0) -> Handle:
    h = _handles.get(device)
    if h is None or h.h is None:
        h = Handle(device)
        _handles[device] = h
    return h


def _dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt not in DTYPES:
        raise TypeError(f"element type {dt} is not a device scalar (f32, f64, i32, i64)")
    return DTYPES[dt]


class DeviceCsr:
    """Device-resident CSR (spam_dcsr): u64 row_ptr, u32 col_idx, T val."""

    def __init__(self, handle: Handle, p: C.c_void_p, keepalive=None):
        self.handle = handle
        self.p = p
        self._keepalive = keepalive

    @staticmethod
    def upload(m: "CsrMatrix", handle: Optional[Handle] = None) -> "DeviceCsr":
        handle = handle or get_handle()
        out = C.c_void_p()
        check(handle.h, handle.L.spam_csr_upload(handle.h, _dtype_code(m.vals.dtype), m.rows_, m.cols_, m.nnz(),
                                                 ptr(m.offsets), ptr(m.indices), ptr(m.vals), C.byref(out)))
        return DeviceCsr(handle, out)

    @staticmethod
    def wrap(handle: Handle, dtype, rows: int, cols: int, nnz: int, d_ptr: int, d_idx: int, d_val: int,
             keepalive=None) -> "DeviceCsr":
        """Non-owning view over caller-managed device memory (e.g. torch tensors)."""
        out = C.c_void_p()
        check(handle.h, handle.L.spam_dcsr_wrap(handle.h, _dtype_code(dtype), rows, cols, nnz, C.c_void_p(d_ptr),
                                                C.c_void_p(d_idx), C.c_void_p(d_val), C.byref(out)))
        return DeviceCsr(handle, out, keepalive)

    @staticmethod
    def from_triplets_sharded(handle: Handle, dtype, rows: int, cols: int, n_local: int, d_rows: int, d_cols: int,
                              d_vals: int) -> Tuple["DeviceCsr", int]:
        """Collective.  This rank's contiguous piece of the triplet stream (device pointers: u64 rows, u64 cols, T
        vals) -> its row block of the CSR matrix (spam_dok_to_csr_sharded).  Returns (block, first global row)."""
        out, r0 = C.c_void_p(), C.c_uint64()
        check(handle.h, handle.L.spam_dok_to_csr_sharded(handle.h, _dtype_code(dtype), rows, cols, n_local,
                                                         C.c_void_p(d_rows), C.c_void_p(d_cols), C.c_void_p(d_vals),
                                                         C.byref(r0), C.byref(out)))
        return DeviceCsr(handle, out), r0.value

    def info(self) -> dict:
        dt, r, c, n = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        dp, di, dv = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(None, self.handle.L.spam_dcsr_info(self.p, C.byref(dt), C.byref(r), C.byref(c), C.byref(n), C.byref(dp),
                                                 C.byref(di), C.byref(dv)))
        return {"dtype": NP_OF[dt.value], "rows": r.value, "cols": c.value, "nnz": n.value, "d_ptr": dp.value,
                "d_idx": di.value, "d_val": dv.value}

    def download(self, is_sorted: bool = True) -> "CsrMatrix":
        i = self.info()
        offsets = np.empty(i["rows"] + 1, dtype=np.uint64)
        indices = np.empty(i["nnz"], dtype=np.uint64)
        vals = np.empty(i["nnz"], dtype=i["dtype"])
        check(self.handle.h, self.handle.L.spam_dcsr_download(self.handle.h, self.p, ptr(offsets), ptr(indices),
                                                              ptr(vals)))
        return CsrMatrix(i["rows"], i["cols"], vals, indices, offsets, is_sorted=is_sorted)

    def matmul(self, rhs: "DeviceCsr", reference_order: bool = False) -> "DeviceCsr":
        """C = self * rhs on the device.  reference_order: rows in the reference's B2 = false (slot) order."""
        out = C.c_void_p()
        if reference_order:
            check(self.handle.h, self.handle.L.spam_spgemm_dev_b2(self.handle.h, self.p, rhs.p, 0, C.byref(out)))
        else:
            check(self.handle.h, self.handle.L.spam_spgemm_dev(self.handle.h, self.p, rhs.p, C.byref(out)))
        return DeviceCsr(self.handle, out)

    def matmul_gathered(self, rhs: "DeviceCsr", row_start: int, total_rows: int, nsub: int = 4,
                        mode: int = -1) -> "DeviceCsr":
        """Collective.  self = this rank's row block of A, rhs = all of B: returns the WHOLE product, assembled on
        every rank (spam_spgemm_gathered).  The result is a view of the handle's gather buffers, valid until the
        next gathered product."""
        out = C.c_void_p()
        check(self.handle.h, self.handle.L.spam_spgemm_gathered(self.handle.h, self.p, rhs.p, row_start, total_rows,
                                                                nsub, mode, C.byref(out)))
        return DeviceCsr(self.handle, out)

    def spmv_gathered(self, d_x: int, d_y_full: int, rows_of) -> None:
        """Collective.  y = A x with self = this rank's row block; d_y_full holds all rows on return."""
        rows_of = np.ascontiguousarray(rows_of, dtype=np.uint64)
        check(self.handle.h, self.handle.L.spam_spmv_gathered(self.handle.h, self.p, C.c_void_p(d_x),
                                                              C.c_void_p(d_y_full), ptr(rows_of)))

    def add(self, rhs: "DeviceCsr") -> "DeviceCsr":
        """impl Add for CsrMatrix (apply_elementwise, lib.rs:83-149) on the device."""
        out = C.c_void_p()
        check(self.handle.h, self.handle.L.spam_dcsr_ewise(self.handle.h, 0, self.p, rhs.p, C.byref(out)))
        return DeviceCsr(self.handle, out)

    def sub(self, rhs: "DeviceCsr") -> "DeviceCsr":
        out = C.c_void_p()
        check(self.handle.h, self.handle.L.spam_dcsr_ewise(self.handle.h, 1, self.p, rhs.p, C.byref(out)))
        return DeviceCsr(self.handle, out)

    def transpose(self) -> "DeviceCsr":
        """Matrix::transpose (spam_csr/src/lib.rs:256-264) on the device; rows of the result are sorted."""
        out = C.c_void_p()
        check(self.handle.h, self.handle.L.spam_dcsr_transpose(self.handle.h, self.p, C.byref(out)))
        return DeviceCsr(self.handle, out)

    def slice_rows(self, r0: int, r1: int) -> "DeviceCsr":
        out = C.c_void_p()
        check(self.handle.h, self.handle.L.spam_dcsr_slice_rows(self.handle.h, self.p, r0, r1, C.byref(out)))
        return DeviceCsr(self.handle, out)

    def select_rows(self, rows) -> "DeviceCsr":
        """The listed rows (any order) as a new device matrix."""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = C.c_void_p()
        check(self.handle.h, self.handle.L.spam_dcsr_select_rows(self.handle.h, self.p, ptr(rows), rows.shape[0],
                                                                 C.byref(out)))
        return DeviceCsr(self.handle, out)

    def rows_to_parts(self, rhs: "DeviceCsr", parts: int, balance: str = "flops") -> Tuple[np.ndarray, int]:
        """Contiguous row blocks by the rows_to_threads formula (mul_hash.rs:51-62), balanced on the raw
        product counts (`flops`, the reference's rule) or on the device-time estimate (`cost`)."""
        if balance not in ("flops", "cost"):
            raise ValueError("balance must be 'flops' or 'cost'")
        starts = np.zeros(parts + 1, dtype=np.uint64)
        total = C.c_uint64()
        fn = self.handle.L.spam_rows_to_parts if balance == "flops" else self.handle.L.spam_rows_to_parts_cost
        check(self.handle.h, fn(self.handle.h, self.p, rhs.p, parts, ptr(starts), C.byref(total)))
        return starts, total.value

    def free(self):
        if self.p:
            self.handle.L.spam_dcsr_free(self.handle.h, self.p)
            self.p = None

    def __del__(self):  # pragma: no cover
        try:
            if self.handle.h:
                self.free()
        except Exception:
            pass


class DokMatrix:
    """Host container with the reference DokMatrix's insertion semantics (spam_dok/src/lib.rs:32-36,
    :167-176).  It is an input format, not a compute path."""

    def __init__(self, rows: int, cols: int, dtype=np.float64):
        if rows <= 0 or cols <= 0:
            raise ValueError("rows and cols are NonZeroUsize")
        self.rows_, self.cols_ = int(rows), int(cols)
        self.dtype = np.dtype(dtype)
        self.entries = {}

    @classmethod
    def new(cls, size: Tuple[int, int], dtype=np.float64) -> "DokMatrix":
        return cls(size[0], size[1], dtype)

    def rows(self) -> int:
        return self.rows_

    def cols(self) -> int:
        return self.cols_

    def nnz(self) -> int:
        return len(self.entries)

    def get_element(self, pos):
        i, j = pos
        if not (0 <= i < self.rows_ and 0 <= j < self.cols_):
            raise IndexError("IndexError")
        return self.entries.get((i, j))

    def set_element(self, pos, t):
        i, j = pos
        if not (0 <= i < self.rows_ and 0 <= j < self.cols_):
            raise IndexError("IndexError")
        t = self.dtype.type(t)
        if t == 0:
            return self.entries.pop((i, j), None)
        old = self.entries.get((i, j))
        self.entries[(i, j)] = t
        return old

    def iter(self):
        for k in sorted(self.entries):
            yield k, self.entries[k]

    def invariants(self) -> bool:
        return all(0 <= r < self.rows_ and 0 <= c < self.cols_ and v != 0 for (r, c), v in self.entries.items())


class CsrMatrix:
    """CSR matrix with the reference's fields (spam_csr/src/lib.rs:25-32).  `is_sorted` is the
    IS_SORTED const generic: rows strictly increasing by column, or merely distinct."""

    def __init__(self, rows: int, cols: int, vals, indices, offsets, is_sorted: bool = True):
        if rows <= 0 or cols <= 0:
            raise ValueError("rows and cols are NonZeroUsize")
        self.rows_, self.cols_ = int(rows), int(cols)
        self.vals = np.ascontiguousarray(vals)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint64)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.is_sorted = bool(is_sorted)
        _dtype_code(self.vals.dtype)

    # ---- Matrix trait (spam_matrix/src/lib.rs:15-27), the parts on the hot path's boundary ----
    @classmethod
    def new(cls, size: Tuple[int, int], dtype=np.float64, is_sorted: bool = True) -> "CsrMatrix":
        r, c = size
        return cls(r, c, np.empty(0, dtype=dtype), np.empty(0, np.uint64), np.zeros(r + 1, np.uint64), is_sorted)

    @classmethod
    def new_square(cls, n: int, dtype=np.float64) -> "CsrMatrix":
        return cls.new((n, n), dtype)

    @classmethod
    def identity(cls, n: int, dtype=np.float64) -> "CsrMatrix":  # lib.rs:177-185
        return cls(n, n, np.ones(n, dtype=dtype), np.arange(n, dtype=np.uint64), np.arange(n + 1, dtype=np.uint64))

    def rows(self) -> int:
        return self.rows_

    def cols(self) -> int:
        return self.cols_

    def nnz(self) -> int:
        return int(self.indices.shape[0])

    def get_element(self, pos):  # lib.rs:199-213
        i, j = pos
        if not (0 <= i < self.rows_ and 0 <= j < self.cols_):
            raise IndexError("IndexError")
        lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
        hit = np.nonzero(self.indices[lo:hi] == j)[0]
        return self.vals[lo + hit[0]] if hit.size else None

    def iter(self) -> Iterator:  # lib.rs:35-45
        for r in range(self.rows_):
            for e in range(int(self.offsets[r]), int(self.offsets[r + 1])):
                yield (r, int(self.indices[e])), self.vals[e]

    def invariants(self) -> bool:  # lib.rs:47-81
        o, idx = self.offsets, self.indices
        if idx.shape[0] != self.vals.shape[0]:            # invariant1
            return False
        if o.shape[0] != self.rows_ + 1:                  # invariant2
            return False
        if np.any(o[1:] < o[:-1]):                        # invariant3
            return False
        if int(o[self.rows_]) != idx.shape[0]:            # invariant4
            return False
        if idx.size and int(idx.max()) >= self.cols_:     # invariant5
            return False
        if int(o[0]) != 0:                                # invariant7
            return False
        if idx.size:                                      # invariant6
            row_of = np.repeat(np.arange(self.rows_, dtype=np.int64), np.diff(o).astype(np.int64))
            if self.is_sorted:
                same = row_of[1:] == row_of[:-1]
                if np.any(same & (idx[1:] <= idx[:-1])):
                    return False
            else:
                key = row_of.astype(np.uint64) * np.uint64(self.cols_) + idx
                if np.unique(key).shape[0] != key.shape[0]:
                    return False
        return True

    # ---- the hot path -----------------------------------------------------------------------
    def mul_hash(self, rhs: "CsrMatrix", sorted_output: bool = False, handle: Optional[Handle] = None,
                 reference_order: bool = False) -> "CsrMatrix":
        """CsrMatrix::mul_hash::<B1, B2> (mul_hash.rs:13-36) on the GPU through the two-phase C ABI
        (spam_spgemm_symbolic / spam_spgemm_numeric).  `sorted_output` is B2.  By default the rows come back
        sorted by column, which satisfies both variants; with B2 = false and `reference_order` they
        come back in the reference's own order (the slot order of its hash map), column for column."""
        if self.vals.dtype != rhs.vals.dtype:
            raise TypeError("operand element types differ")
        handle = handle or get_handle()
        L, h = handle.L, handle.h
        c_ptr = np.empty(self.rows_ + 1, dtype=np.uint64)
        nnz = C.c_uint64()
        same = rhs is self
        check(h, L.spam_spgemm_symbolic(h, _dtype_code(self.vals.dtype), self.rows_, self.cols_, ptr(self.offsets),
                                        ptr(self.indices), ptr(self.vals), rhs.rows_, rhs.cols_,
                                        ptr(self.offsets if same else rhs.offsets),
                                        ptr(self.indices if same else rhs.indices),
                                        ptr(self.vals if same else rhs.vals), ptr(c_ptr), C.byref(nnz)))
        c_idx = np.empty(nnz.value, dtype=np.uint64)   # Vec::with_capacity(nnz), mul_hash.rs:119
        c_val = np.empty(nnz.value, dtype=self.vals.dtype)
        check(h, L.spam_spgemm_numeric(h, ptr(c_idx), ptr(c_val), 0 if (reference_order and not sorted_output) else 1))
        return CsrMatrix(self.rows_, rhs.cols_, c_val, c_idx, c_ptr, is_sorted=sorted_output)

    def __mul__(self, rhs: "CsrMatrix") -> "CsrMatrix":  # impl Mul for &CsrMatrix, lib.rs:292-297
        return self.mul_hash(rhs, sorted_output=False)

    __matmul__ = __mul__

    # ---- elementwise add / sub (impl Add / Sub -> apply_elementwise, lib.rs:83-149, 276-290) ----
    def _ewise(self, rhs: "CsrMatrix", op: int, handle: Optional[Handle]) -> "CsrMatrix":
        if self.vals.dtype != rhs.vals.dtype:
            raise TypeError("operand element types differ")
        if (self.rows_, self.cols_) != (rhs.rows_, rhs.cols_):
            raise _lib.DimensionMismatch(2, "matrices must have identical dimensions")   # assert_eq!, lib.rs:87-91
        handle = handle or get_handle()
        L, h = handle.L, handle.h
        c_ptr = np.empty(self.rows_ + 1, dtype=np.uint64)
        nnz = C.c_uint64()
        same = rhs is self
        if not self.is_sorted:
            op |= 2    # IS_SORTED = false: entries only in the left operand are kept untouched (lib.rs:119-137)
        check(h, L.spam_csr_ewise(h, op, _dtype_code(self.vals.dtype), self.rows_, self.cols_, ptr(self.offsets),
                                  ptr(self.indices), ptr(self.vals), ptr(self.offsets if same else rhs.offsets),
                                  ptr(self.indices if same else rhs.indices), ptr(self.vals if same else rhs.vals),
                                  ptr(c_ptr), C.byref(nnz)))
        c_idx = np.empty(nnz.value, dtype=np.uint64)
        c_val = np.empty(nnz.value, dtype=self.vals.dtype)
        check(h, L.spam_csr_ewise_fetch(h, ptr(c_idx), ptr(c_val)))
        return CsrMatrix(self.rows_, self.cols_, c_val, c_idx, c_ptr, is_sorted=self.is_sorted)

    def add(self, rhs: "CsrMatrix", handle: Optional[Handle] = None) -> "CsrMatrix":
        return self._ewise(rhs, 0, handle)

    def sub(self, rhs: "CsrMatrix", handle: Optional[Handle] = None) -> "CsrMatrix":
        return self._ewise(rhs, 1, handle)

    def __add__(self, rhs: "CsrMatrix") -> "CsrMatrix":
        return self._ewise(rhs, 0, None)

    def __sub__(self, rhs: "CsrMatrix") -> "CsrMatrix":
        return self._ewise(rhs, 1, None)

    def transpose(self, handle: Optional[Handle] = None) -> "CsrMatrix":
        """Matrix::transpose (spam_csr/src/lib.rs:256-264): every stored entry (i, j, v), explicit zeros
        included, becomes (j, i, v); the result's rows are sorted by column."""
        handle = handle or get_handle()
        nnz = self.nnz()
        t_ptr = np.empty(self.cols_ + 1, dtype=np.uint64)
        t_idx = np.empty(nnz, dtype=np.uint64)
        t_val = np.empty(nnz, dtype=self.vals.dtype)
        check(handle.h, handle.L.spam_csr_transpose(handle.h, _dtype_code(self.vals.dtype), self.rows_, self.cols_,
                                                    ptr(self.offsets), ptr(self.indices), ptr(self.vals), ptr(t_ptr),
                                                    ptr(t_idx), ptr(t_val)))
        return CsrMatrix(self.cols_, self.rows_, t_val, t_idx, t_ptr, is_sorted=True)

    def spmv(self, x, handle: Optional[Handle] = None) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=self.vals.dtype)
        if x.shape != (self.cols_,):
            raise _lib.DimensionMismatch(2, f"x has shape {x.shape}, expected ({self.cols_},)")
        handle = handle or get_handle()
        y = np.empty(self.rows_, dtype=self.vals.dtype)
        check(handle.h, handle.L.spam_spmv(handle.h, _dtype_code(self.vals.dtype), self.rows_, self.cols_,
                                           ptr(self.offsets), ptr(self.indices), ptr(self.vals), ptr(x), ptr(y)))
        return y

    # ---- DOK -> CSR -------------------------------------------------------------------------
    @classmethod
    def from_triplets(cls, rows: int, cols: int, tri_rows, tri_cols, tri_vals,
                      handle: Optional[Handle] = None) -> "CsrMatrix":
        """DokMatrix::new((rows, cols)), one set_element per triplet in stream order, then
        CsrMatrix::from(dok): last write wins, writing zero deletes; result rows sorted."""
        tri_vals = np.ascontiguousarray(tri_vals)
        tri_rows = np.ascontiguousarray(tri_rows, dtype=np.uint64)
        tri_cols = np.ascontiguousarray(tri_cols, dtype=np.uint64)
        if not (tri_rows.shape == tri_cols.shape == tri_vals.shape):
            raise ValueError("triplet arrays differ in length")
        handle = handle or get_handle()
        c_ptr = np.empty(rows + 1, dtype=np.uint64)
        nnz = C.c_uint64()
        check(handle.h, handle.L.spam_dok_to_csr(handle.h, _dtype_code(tri_vals.dtype), rows, cols, tri_vals.shape[0],
                                                 ptr(tri_rows), ptr(tri_cols), ptr(tri_vals), ptr(c_ptr),
                                                 C.byref(nnz)))
        c_idx = np.empty(nnz.value, dtype=np.uint64)
        c_val = np.empty(nnz.value, dtype=tri_vals.dtype)
        check(handle.h, handle.L.spam_dok_to_csr_fetch(handle.h, ptr(c_idx), ptr(c_val)))
        return cls(rows, cols, c_val, c_idx, c_ptr, is_sorted=True)

    @classmethod
    def from_dok(cls, dok: DokMatrix, handle: Optional[Handle] = None) -> "CsrMatrix":
        """impl From<DokMatrix<T>> for CsrMatrix<T, true> (lib.rs:315-334)."""
        n = len(dok.entries)
        r = np.fromiter((k[0] for k in dok.entries), dtype=np.uint64, count=n)
        c = np.fromiter((k[1] for k in dok.entries), dtype=np.uint64, count=n)
        v = np.fromiter(dok.entries.values(), dtype=dok.dtype, count=n)
        return cls.from_triplets(dok.rows_, dok.cols_, r, c, v, handle)

    def to_dok(self) -> DokMatrix:
        """impl From<CsrMatrix> for DokMatrix (spam_csr/src/lib.rs:360-384): zeros are dropped by set_element."""
        d = DokMatrix(self.rows_, self.cols_, self.vals.dtype)
        for pos, t in self.iter():
            d.set_element(pos, t)
        return d

    def __repr__(self):
        return (f"CsrMatrix<{self.vals.dtype}, {str(self.is_sorted).lower()}>"
                f"({self.rows_}x{self.cols_}, nnz={self.nnz()})")
