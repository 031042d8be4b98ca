"""MatrixMarket ingest: text -> triplets (host, libspam_cuda.so's spam_mm_parse) -> device
DOK -> CSR.  Mirrors `parse_matrix_market` / `MatrixType` (spam_dok/src/lib.rs:282-478) and the writer
`into_float_matrix_market` (:480-489); the reference's bench reads its inputs this way
(spam_csr/src/lib.rs:419-431).  The parser is host code: it needs the built library but no GPU."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib


class SpamMM(C.Structure):
    _fields_ = [("kind", C.c_int), ("rows", C.c_uint64), ("cols", C.c_uint64), ("declared_entries", C.c_uint64),
                ("n", C.c_uint64), ("tri_rows", C.c_void_p), ("tri_cols", C.c_void_p), ("tri_vals", C.c_void_p),
                ("err", C.c_char * 96)]


class FromMatrixMarketError(ValueError):
    """FromMatrixMarketError (spam_dok/src/lib.rs:268-280) and the reference's todo!() cases."""


def parse_matrix_market(text) -> Tuple[str, int, int, np.ndarray, np.ndarray, np.ndarray]:
    """-> (kind, rows, cols, tri_rows, tri_cols, tri_vals); kind 'integer' (i64 values) or 'real' (f64): the
    entries of MatrixType::Integer / MatrixType::Real as a stream for DokMatrix.set_element / from_triplets
    (zeros already skipped, symmetric entries mirrored, 0-based indices, a later duplicate replaces)."""
    L = _lib.load()
    data = text.encode() if isinstance(text, str) else bytes(text)
    m = SpamMM()
    L.spam_mm_parse.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(SpamMM)]
    L.spam_mm_free.argtypes = [C.POINTER(SpamMM)]
    L.spam_mm_free.restype = None
    st = L.spam_mm_parse(data, len(data), C.byref(m))
    if st != 0:
        msg = m.err.decode(errors="replace")
        if st == 8:
            raise IndexError(f"IndexError: {msg}")
        raise FromMatrixMarketError(f"{_lib.STATUS_NAMES.get(st, st)}: {msg}")
    try:
        n = int(m.n)
        dt = np.int64 if m.kind == _lib.DTYPES[np.dtype(np.int64)] else np.float64

        def take(p, dtype):
            if n == 0:
                return np.empty(0, dtype)
            return np.frombuffer((C.c_char * (n * 8)).from_address(p), dtype=dtype, count=n).copy()
        return ("integer" if dt is np.int64 else "real", int(m.rows), int(m.cols), take(m.tri_rows, np.uint64),
                take(m.tri_cols, np.uint64), take(m.tri_vals, dt))
    finally:
        L.spam_mm_free(C.byref(m))


def load_matrix_market(path: str, handle=None):
    """File -> CsrMatrix<T, true> through the device DOK -> CSR build (CsrMatrix::from(dok), lib.rs:315-334)."""
    from .csr import CsrMatrix
    with open(path, "rb") as f:
        kind, rows, cols, tr, tc, tv = parse_matrix_market(f.read())
    return CsrMatrix.from_triplets(rows, cols, tr, tc, tv, handle=handle)


def into_float_matrix_market(m) -> str:
    """`into_float_matrix_market` (spam_dok/src/lib.rs:480-489): real general, 1-based, entries in (row, col)
    order, floats in Rust's `{}` Display form (shortest round-trip digits, never an exponent)."""
    lines = ["%%MatrixMarket matrix coordinate real general", f"{m.rows()} {m.cols()} {m.nnz()}"]
    for (i, j), t in m.iter():
        t = float(t)
        if t != t:
            s = "NaN"
        elif t in (float("inf"), float("-inf")):
            s = "inf" if t > 0 else "-inf"
        else:
            s = np.format_float_positional(t, unique=True, trim="-")
        lines.append(f"{i + 1} {j + 1} {s}")
    return "\n".join(lines) + "\n"
