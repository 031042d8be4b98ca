"""GPU parity at BASELINE.json's FULL sizes for the configurations the reduced tests only sample (-m gpu):
C3 (27-point 160^3), C4 (R-MAT scale 22) and C5 (1M x 4M, i64, A*A^T + DOK->CSR).  C1 and C2 at full size live in
test_gpu_parity.py.  north_star: "bit-exact structure and in-tolerance values versus spam_csr on every config".

C3 and C5 are compared entry by entry with the oracle.  C4's product has 2.53e9 entries (40 GB as host u64/f64):
its row_ptr is compared in full with the oracle's symbolic pass, and col_idx / values on the heaviest rows (all of
the global-table / column-range bin's biggest rows) plus a sample of row blocks spread over the matrix, pulled
off the device with spam_dcsr_select_rows and compared with the oracle run on the same rows of A.
"""
import numpy as np
import pytest

import sparse_matrix_b200 as S
from sparse_matrix_b200 import generators as G
from util import TOL, as_csr_matrix, check_against_oracle

pytestmark = pytest.mark.gpu

HEAVY = S._lib.HEAVY_BIN


def _rows_of(mat, rows):
    """The listed rows of a host CSR tuple as a new tuple."""
    r, c, off, idx, val = mat
    rows = np.asarray(rows, dtype=np.int64)
    lo, hi = off[rows].astype(np.int64), off[rows + 1].astype(np.int64)
    lens = hi - lo
    noff = np.zeros(len(rows) + 1, np.uint64)
    noff[1:] = np.cumsum(lens)
    take = np.repeat(lo - noff[:-1].astype(np.int64), lens) + np.arange(int(lens.sum()), dtype=np.int64)
    return len(rows), c, noff, idx[take], val[take]


def _compare_rows(oracle, a_rows, b, got, exact=False):
    off, idx, val = oracle.mul_hash(a_rows, b, True)
    assert np.array_equal(got.offsets, off), "row_ptr of the row sample differs"
    assert np.array_equal(got.indices, idx), "col_idx of the row sample differs"
    if exact or np.dtype(val.dtype).kind != "f":
        assert np.array_equal(got.vals, val)
    else:
        _, _, sabs = oracle.mul_hash(a_rows[:4] + (np.abs(a_rows[4]),), b[:4] + (np.abs(b[4]),), True)
        err = np.abs(got.vals - val)
        assert np.all(err <= TOL[np.dtype(val.dtype)] * sabs), float(np.max(err / np.maximum(sabs, 1e-300)))


def test_config3_stencil27_full_size(oracle, handle):
    """27-point stencil on 160^3: 4 096 000 rows, 109 215 352 entries, 2 924 207 000 products, nnz(C) = 794^3."""
    s = G.stencil27(160)
    assert s[0] == 4_096_000 and len(s[3]) == 478 ** 3
    dA = S.DeviceCsr.upload(as_csr_matrix(s), handle)
    handle.set_timing(True)
    dC = dA.matmul(dA)
    st = handle.stats()
    handle.set_timing(False)
    assert st["flops"] == 1430 ** 3 and st["nnz_c"] == 794 ** 3
    c = dC.download()
    dC.free(); dA.free()
    check_against_oracle(oracle, s, s, c)


def test_config4_rmat_full_size(oracle, handle):
    """R-MAT(0.45, 0.15, 0.15, 0.25) scale 22, edge factor 16, seed 42: u64 row_ptr past 2^31, every bin in use."""
    r = G.rmat(22)
    flops, per_row = G.spgemm_counts(r, r)
    dA = S.DeviceCsr.upload(as_csr_matrix(r), handle)
    handle.set_timing(True)
    dC = dA.matmul(dA)
    st = handle.stats()
    handle.set_timing(False)
    assert st["flops"] == flops and st["nnz_c"] > 2 ** 31
    assert st["sym_bin_rows"][HEAVY] > 0, st["sym_bin_rows"]
    assert st["num_bin_rows"][HEAVY] + st["num_bin_rows"][15] > 0, st["num_bin_rows"]   # rows past the shared-memory bins
    # row_ptr, all 4 194 305 entries, against the reference's symbolic pass
    z = oracle.symbolic(r[0], r[2], r[3], r[2], r[3])
    want_ptr = np.zeros(r[0] + 1, np.uint64)
    np.cumsum(z, out=want_ptr[1:])
    i = dC.info()
    got_ptr = np.empty(r[0] + 1, np.uint64)
    S._lib.check(handle.h, handle.L.spam_dcsr_download(handle.h, dC.p, S._lib.ptr(got_ptr), None, None))
    assert i["nnz"] == int(want_ptr[-1]) == st["nnz_c"]
    assert np.array_equal(got_ptr, want_ptr)
    # col_idx / values: the 2000 heaviest rows (every row above 16 384 products), the first
    # 2048 rows (R-MAT's hubs), and 48 blocks of 1024 rows spread over the matrix (about 1.2 % of the rows)
    heavy = np.argsort(per_row)[-2000:]
    assert per_row[heavy].min() > 8192 and (per_row > 16384).sum() <= 2000
    rng = np.random.default_rng(22)
    starts = rng.integers(0, r[0] - 1024, size=48)
    sample = np.unique(np.concatenate([heavy, np.arange(2048)] + [np.arange(s0, s0 + 1024) for s0 in starts]))
    for part in np.array_split(sample, 8):          # bounded host memory per oracle call
        sub = dC.select_rows(part)
        got = sub.download()
        sub.free()
        _compare_rows(oracle, _rows_of(r, part), r, got)
    dC.free(); dA.free()


def test_config5_rectangular_full_size(oracle, handle):
    """1M x 4M, 8 per row, i64: A * A^T with A^T made by the device transpose, and the DOK -> CSR build of the
    same matrix from a shuffled triplet stream with rewrites and deletions; everything bit-exact."""
    a = G.uniform_random(1_000_000, 4_000_000, 8, seed=5, dtype=np.int64, int_range=1 << 15)
    A = as_csr_matrix(a)
    dA = S.DeviceCsr.upload(A, handle)
    dT = dA.transpose()
    t = dT.download()
    want_t = oracle.transpose(a)
    assert np.array_equal(t.offsets, want_t[0]) and np.array_equal(t.indices, want_t[1]) and np.array_equal(t.vals, want_t[2])
    dC = dA.matmul(dT)
    c = dC.download()
    for d in (dC, dT, dA):
        d.free()
    at = (a[1], a[0]) + want_t
    check_against_oracle(oracle, a, at, c)
    tr, tc, tv = G.triplets_with_rewrites(a, seed=5)
    got = S.CsrMatrix.from_triplets(a[0], a[1], tr, tc, tv, handle=handle)
    off, idx, val = oracle.dok_to_csr(a[0], a[1], tr, tc, tv)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
