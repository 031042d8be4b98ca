"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> libspam_cuda.so),
against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): row_ptr and col_idx bit-exact against mul_hash::<_, true>; integer
values bit-exact; f64 / f32 values within 1e-12 / 1e-5 of each entry's sum of |products|.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import sparse_matrix_b200 as S
from sparse_matrix_b200 import distributed as D
from sparse_matrix_b200 import generators as G
from util import TOL, as_csr_matrix, check_against_oracle, random_csr

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "linprobe_kat.json")))
ALL_DTYPES = [np.float64, np.float32, np.int32, np.int64]
MERGE, HEAVY = S._lib.MERGE_BIN, S._lib.HEAVY_BIN


def gpu_mul(a, b, handle, sorted_output=True):
    A = as_csr_matrix(a, is_sorted=False)
    B = A if b is a else as_csr_matrix(b, is_sorted=False)
    return A.mul_hash(B, sorted_output=sorted_output, handle=handle)


# ---------------------------------------------------------------------------------------------
# known answers and the reference's own property test
# ---------------------------------------------------------------------------------------------
def test_golden_kats_on_gpu(handle):
    g = GOLD["map_slot_order"]
    a = (1, 4, g["a"]["offsets"], g["a"]["indices"], np.array(g["a"]["vals"]))
    b = (4, 4, g["b"]["offsets"], g["b"]["indices"], np.array(g["b"]["vals"]))
    c = gpu_mul(a, b, handle)
    assert c.offsets.tolist() == g["sorted"]["offsets"] and c.indices.tolist() == g["sorted"]["indices"]
    assert c.vals.tolist() == g["sorted"]["vals"]
    # hand-derived slot orders: one map re-sized per row, and a collision chain with wrap-around and accumulation
    for name in ("map_reuse_shrink", "map_collision_chain"):
        g = GOLD[name]
        a = (g["a"]["rows"], g["a"]["cols"], g["a"]["offsets"], g["a"]["indices"], np.array(g["a"]["vals"]))
        b = ((65, 65, np.arange(66), np.arange(65), np.ones(65)) if g["b"] == "identity 65" else
             (g["b"]["rows"], g["b"]["cols"], g["b"]["offsets"], g["b"]["indices"], np.array(g["b"]["vals"])))
        c = gpu_mul(a, b, handle)
        assert c.indices.tolist() == g["sorted"]["indices"] and c.vals.tolist() == g["sorted"]["vals"]
        c = as_csr_matrix(a, False).mul_hash(as_csr_matrix(b, False), sorted_output=False, handle=handle, reference_order=True)
        assert c.offsets.tolist() == g["unsorted"]["offsets"] and c.indices.tolist() == g["unsorted"]["indices"], name
        assert c.vals.tolist() == g["unsorted"]["vals"]
    g = GOLD["unfused_cancellation"]
    av = np.array([float.fromhex(x) for x in g["a"]["vals_hex"]])
    bv = np.array([float.fromhex(x) for x in g["b"]["vals_hex"]])
    c = gpu_mul((1, 2, g["a"]["offsets"], g["a"]["indices"], av), (2, 8, g["b"]["offsets"], g["b"]["indices"], bv),
                handle)
    assert c.offsets.tolist() == [0, 1] and c.indices.tolist() == [5]
    assert c.vals.tolist() == [0.0]          # no FMA on the device either, and the zero is kept


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_reference_property_dense_dok_equivalence(oracle, handle, dtype):
    """spam_csr/src/tests.rs:356-371 with a device scalar in place of Wrapping<i8>: dims <= 4,
    shuffled rows, wrapping arithmetic; compare through the dense DOK product (zero-insensitive)."""
    rng = np.random.default_rng(1234)
    info = np.iinfo(dtype)
    for _ in range(150):
        l, m, n = (int(x) for x in rng.integers(1, 5, size=3))
        da = np.zeros((l, m), dtype)
        db = np.zeros((m, n), dtype)
        for d in (da, db):
            for _k in range(int(rng.integers(0, 2 * d.size + 1))):
                d[rng.integers(0, d.shape[0]), rng.integers(0, d.shape[1])] = rng.integers(info.min, info.max,
                                                                                            dtype=dtype)
        a = _dense_to_unsorted(da, rng)
        b = _dense_to_unsorted(db, rng)
        c = gpu_mul(a, b, handle, sorted_output=False)
        assert c.invariants()
        got = np.zeros((l, n), dtype)
        for (r, col), t in c.iter():
            got[r, col] = t
        assert np.array_equal(got, oracle.dok_dense_mul(da, db))


def _dense_to_unsorted(dense, rng):
    rows, cols = dense.shape
    off, idx, val = [0], [], []
    for r in range(rows):
        c = rng.permutation(np.nonzero(dense[r])[0])
        idx.extend(c.tolist())
        val.extend(dense[r, c].tolist())
        off.append(len(idx))
    return rows, cols, np.array(off, np.uint64), np.array(idx, np.uint64), np.array(val, dense.dtype)


# ---------------------------------------------------------------------------------------------
# seeded random products against the oracle, every dtype
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 9), (300, 200, 400), (2000, 2000, 2000), (50, 3000, 60)])
def test_random_products(oracle, handle, dtype, shape):
    l, m, n = shape
    rng = np.random.default_rng(hash((l, m, n)) % 2**32)
    a = random_csr(rng, l, m, rng.integers(0, min(m, 24) + 1, size=l), dtype=dtype, sorted_rows=False)
    b = random_csr(rng, m, n, rng.integers(0, min(n, 24) + 1, size=m), dtype=dtype, sorted_rows=False)
    c = gpu_mul(a, b, handle)
    check_against_oracle(oracle, a, b, c)


@pytest.mark.parametrize("dtype", [np.float64, np.int64])
def test_every_bin_is_exercised(oracle, handle_nosort, handle_esc, dtype):
    """Rows whose flop / nnz land in each symbolic and numeric bin: hash bins 1..8 and the global-table bin 9 by
    default; with SPAM_ESC=2 the rows that do not compress (every product a new column) take the bucket-sort bins
    11..15 instead."""
    handle = handle_nosort
    rng = np.random.default_rng(99)
    inner, n = 3000, 60000
    # (1) A row i has `deg[i]` entries; B rows have 40 entries => flop = 40*deg, nnz close to it
    deg = np.array([0, 1, 2, 5, 8, 20, 40, 90, 150, 300, 420, 700, 1000, 1500, 2500] + [3] * 40 + [60] * 20)
    a = random_csr(rng, len(deg), inner, deg, dtype=dtype, sorted_rows=False)
    b = random_csr(rng, inner, n, 40, dtype=dtype, sorted_rows=False)
    handle.set_timing(True)
    c = gpu_mul(a, b, handle)
    st = handle.stats()
    assert all(x > 0 for x in st["sym_bin_rows"][:10]), st["sym_bin_rows"]
    assert all(x > 0 for x in st["num_bin_rows"][:10]) and sum(st["num_bin_rows"][11:16]) == 0, st["num_bin_rows"]
    off, idx, val = check_against_oracle(oracle, a, b, c)
    assert st["nnz_c"] == len(idx) and st["flops"] == G.spgemm_counts(a, b)[0] and st["kernel_launches"] >= 10
    c = gpu_mul(a, b, handle_esc)
    st = handle_esc.stats()
    assert all(x > 0 for x in st["num_bin_rows"][:4]) and all(x > 0 for x in st["num_bin_rows"][11:16]), st["num_bin_rows"]
    assert st["fallbacks"][3] == 0          # uniform columns: no crowded bucket
    check_against_oracle(oracle, a, b, c)
    # (2) the same B three times over (rows k, k + inner, k + 2 inner hold the same columns) and A rows that
    # reference all three copies: every column is produced three times, so the rows compress and are hashed
    deg3 = np.array([0, 1, 2, 4, 7, 12, 25, 50, 100, 200, 400, 2500] + [3] * 40 + [60] * 20)
    a1 = random_csr(rng, len(deg3), inner, deg3, dtype=dtype, sorted_rows=False)
    lens = np.diff(a1[2]).astype(np.int64)
    idx3 = np.concatenate([np.concatenate([a1[3][int(lo):int(hi)] + np.uint64(k * inner) for k in range(3)])
                           for lo, hi in zip(a1[2][:-1], a1[2][1:])]) if len(a1[3]) else a1[3]
    off3 = np.zeros(len(deg3) + 1, np.uint64)
    off3[1:] = np.cumsum(3 * lens)
    v3 = rng.integers(1, 9, size=len(idx3)).astype(dtype)
    a3 = (len(deg3), 3 * inner, off3, idx3, v3)
    b3 = (3 * inner, n, np.concatenate([b[2][:-1] + np.uint64(k * len(b[3])) for k in range(3)] + [[np.uint64(3 * len(b[3]))]]),
          np.tile(b[3], 3), rng.permutation(np.tile(b[4], 3)))
    c3 = gpu_mul(a3, b3, handle)
    st = handle.stats()
    handle.set_timing(False)
    assert all(x > 0 for x in st["sym_bin_rows"][:10]), st["sym_bin_rows"]
    assert all(x > 0 for x in st["num_bin_rows"][:10]), st["num_bin_rows"]
    check_against_oracle(oracle, a3, b3, c3)
    # many duplicates: few distinct columns but a large flop count (numeric bin chosen by nnz, not flop)
    b2 = random_csr(rng, inner, 24, 20, dtype=dtype, sorted_rows=False)
    c2 = gpu_mul(a, b2, handle)
    check_against_oracle(oracle, a, b2, c2)


def test_tiny_rows_are_bit_identical_for_floats(oracle, handle, handle_nosort):
    """The thread-per-row path visits products in the reference's order with unfused mul/add, so the
    f64 sums must equal the oracle's bit for bit, not just within tolerance."""
    for dtype in (np.float64, np.float32):
        p = G.poisson2d(96, dtype=dtype)
        rng = np.random.default_rng(5)
        p = p[:4] + (rng.uniform(-1, 1, size=p[4].shape).astype(dtype),)
        handle.set_timing(True)
        c = gpu_mul(p, p, handle)                       # sorted rows: the merge bin
        st = handle.stats()
        assert st["sym_bin_rows"][MERGE] == p[0] and st["num_bin_rows"][MERGE] == p[0]
        check_against_oracle(oracle, p, p, c, exact_values=True)
        # same matrix with every row shuffled (CsrMatrix<T,false>): the private-hash-table bin
        off, idx, val = p[2], p[3].copy(), p[4].copy()
        for r in range(p[0]):
            lo, hi = int(off[r]), int(off[r + 1])
            perm = rng.permutation(hi - lo)
            idx[lo:hi] = idx[lo:hi][perm]
            val[lo:hi] = val[lo:hi][perm]
        q = (p[0], p[1], off, idx, val)
        c = gpu_mul(q, q, handle_nosort)
        st = handle_nosort.stats()
        handle.set_timing(False)
        assert st["sym_bin_rows"][0] == p[0] and st["num_bin_rows"][0] == p[0]
        check_against_oracle(oracle, q, q, c, exact_values=True)
        # default handle: the unsorted right-hand side is replaced by its cached sorted copy, so the merge bin runs —
        # and the sums are still the reference's bit for bit (one product per entry of A's row, in A's order)
        c = gpu_mul(q, q, handle)
        st = handle.stats()
        assert st["num_bin_rows"][MERGE] == p[0]
        check_against_oracle(oracle, q, q, c, exact_values=True)


def test_merge_bin_mixed_with_other_bins(oracle, handle):
    """Sorted B, A rows of every length: short rows take the merge bin, the rest the hash bins;
    also K = 4 / 6 / 8 head variants and rows with empty B rows among the runs."""
    rng = np.random.default_rng(77)
    for kmax, dtype in ((3, np.float64), (6, np.int64), (8, np.float32), (40, np.float64)):
        rows, inner, n = 5000, 4000, 6000
        bdeg = rng.integers(0, 9, size=inner)
        bdeg[rng.random(inner) < 0.2] = 0
        b = random_csr(rng, inner, n, bdeg, dtype=dtype, sorted_rows=True)
        a = random_csr(rng, rows, inner, rng.integers(0, kmax + 1, size=rows), dtype=dtype, sorted_rows=False)
        handle.set_timing(True)
        c = gpu_mul(a, b, handle)
        st = handle.stats()
        handle.set_timing(False)
        assert st["sym_bin_rows"][MERGE] > 0 and st["num_bin_rows"][MERGE] > 0
        if kmax == 40:
            assert sum(st["sym_bin_rows"][1:9]) > 0 and sum(st["num_bin_rows"][1:9]) > 0
        check_against_oracle(oracle, a, b, c)


def test_merge_bin_window_kernels(oracle, handle_win):
    """SPAM_MERGE_WIN=3: banded matrices (the window of B fits shared memory and is fetched by bulk copies), scattered
    ones (it does not: the block falls back to the global-memory loop) and arrays whose length is not a multiple of 4
    entries (the last entries cannot be part of a 16-byte copy).  Floats bit-identical, as in the default kernels."""
    h = handle_win
    rng = np.random.default_rng(404)
    for n, dtype in ((96, np.float64), (61, np.float32), (50, np.int64), (33, np.int32)):
        p = G.poisson2d(n, dtype=dtype)
        vals = (rng.uniform(-1, 1, size=p[4].shape) if np.issubdtype(dtype, np.floating)
                else rng.integers(-9, 10, size=p[4].shape)).astype(dtype)
        p = p[:4] + (vals,)
        h.set_timing(True)
        c = gpu_mul(p, p, h)
        st = h.stats()
        h.set_timing(False)
        assert st["sym_bin_rows"][MERGE] == p[0] and st["num_bin_rows"][MERGE] == p[0]
        check_against_oracle(oracle, p, p, c, exact_values=True)
    for trial in range(4):
        # nnz(B) = 4q + trial: every tail length; a band of A so that the last rows of B are in some block's window
        inner = 3000
        bdeg = np.full(inner, 4)
        bdeg[-1] = 12
        b = random_csr(rng, inner, 5000, bdeg, dtype=np.float64, sorted_rows=True)
        drop = (b[3].shape[0] - trial) % 4              # shorten the last row to the wanted remainder
        boff = b[2].copy()
        boff[-1] -= np.uint64(drop)
        b = (b[0], b[1], boff, b[3][:b[3].shape[0] - drop], b[4][:b[4].shape[0] - drop])
        assert b[3].shape[0] % 4 == trial and int(boff[-1]) > int(boff[-2])
        rows = inner
        off = np.arange(rows + 1, dtype=np.uint64) * 3
        idx = np.clip(np.arange(rows)[:, None] + np.array([-1, 0, 1])[None, :], 0, inner - 1)
        idx[0] = [0, 1, 2]
        idx[-1] = [inner - 3, inner - 2, inner - 1]
        a = (rows, inner, off, idx.reshape(-1).astype(np.uint64), rng.uniform(-1, 1, size=rows * 3))
        c = gpu_mul(a, b, h)
        check_against_oracle(oracle, a, b, c, exact_values=True)
    # scattered rows of A: windows far larger than the staging buffer
    b = random_csr(rng, 200000, 200000, 6, dtype=np.float64, sorted_rows=True)
    a = random_csr(rng, 20000, 200000, rng.integers(0, 9, size=20000), dtype=np.float64, sorted_rows=True)
    check_against_oracle(oracle, a, b, gpu_mul(a, b, h), exact_values=True)


@pytest.mark.parametrize("pf", [1, 6])
def test_merge_bin_prefetch_kernels(oracle, pf):
    """SPAM_MERGE_PF: the merge kernels that load a head's value with its column (1) and keep the next column of every
    run in a register (2 | 4) — chosen automatically only when A's rows scatter over more of B than the caches hold,
    which no small test matrix does.  Same products in the same order: floats bit-identical."""
    os.environ["SPAM_MERGE_PF"] = str(pf)
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_MERGE_PF"]
    try:
        rng = np.random.default_rng(500 + pf)
        p = G.poisson2d(70, dtype=np.float64)
        p = p[:4] + (rng.uniform(-1, 1, size=p[4].shape),)
        check_against_oracle(oracle, p, p, gpu_mul(p, p, h), exact_values=True)
        for kmax, dtype in ((4, np.float32), (6, np.int64), (8, np.float64)):
            bdeg = rng.integers(0, 12, size=3000)
            bdeg[rng.random(3000) < 0.2] = 0                       # empty runs among the heads
            b = random_csr(rng, 3000, 5000, bdeg, dtype=dtype, sorted_rows=True)
            a = random_csr(rng, 4000, 3000, rng.integers(0, kmax + 1, size=4000), dtype=dtype, sorted_rows=False)
            h.set_timing(True)
            c = gpu_mul(a, b, h)
            st = h.stats()
            h.set_timing(False)
            assert st["num_bin_rows"][MERGE] > 0
            check_against_oracle(oracle, a, b, c, exact_values=True)
    finally:
        h.close()


def test_onepass_product(oracle):
    """SPAM_ONEPASS=1: device-resident products whose rows are all merge rows by the cached statistics (longest row of A
    <= 8, x longest row of B <= 128) run as ONE kernel: symbolic merge, look-back scan over the blocks and numeric merge
    fused; C allocated for the bound nnz(A) x longest row of B.  Bit-identical, also for floats."""
    os.environ["SPAM_ONEPASS"] = "1"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_ONEPASS"]
    try:
        rng = np.random.default_rng(610)
        cases = []
        for n, dtype in ((130, np.float64), (77, np.float32), (40, np.int64)):
            p = G.poisson2d(n, dtype=dtype)
            vals = (rng.uniform(-1, 1, size=p[4].shape) if np.dtype(dtype).kind == "f"
                    else rng.integers(-9, 10, size=p[4].shape)).astype(dtype)
            p = p[:4] + (vals,)
            cases.append((p, p, True))
        bdeg = rng.integers(0, 13, size=3000)
        bdeg[rng.random(3000) < 0.3] = 0
        b = random_csr(rng, 3000, 5000, bdeg, dtype=np.float64)
        a = random_csr(rng, 70_000, 3000, rng.integers(0, 9, size=70_000), dtype=np.float64, sorted_rows=False)
        cases.append((a, b, True))                                   # 547 blocks: several look-back windows
        a2 = random_csr(rng, 500, 3000, rng.integers(0, 12, size=500), dtype=np.float64)
        cases.append((a2, b, False))                                 # longest row of A > 8: two-phase pipeline
        for a, b, one in cases:
            dA = S.DeviceCsr.upload(as_csr_matrix(a, False), h)
            dB = dA if b is a else S.DeviceCsr.upload(as_csr_matrix(b, False), h)
            dC = dA.matmul(dB)
            st = h.stats()
            assert st["fallbacks"][5] == (1 if one else 0), st["fallbacks"]
            if one:
                assert st["num_bin_rows"][MERGE] == a[0]
            check_against_oracle(oracle, a, b, dC.download(), exact_values=one)
            dC.free()
            if dB is not dA:
                dB.free()
            dA.free()
    finally:
        h.close()


def test_edge_cases(oracle, handle):
    # all-empty operands
    a = S.CsrMatrix.new((5, 7))
    b = S.CsrMatrix.new((7, 3))
    c = a.mul_hash(b, sorted_output=True, handle=handle)
    assert c.nnz() == 0 and c.offsets.tolist() == [0] * 6 and c.invariants() and (c.rows(), c.cols()) == (5, 3)
    # empty rows in A and empty rows in B that A points at (flop 0 rows, mul_hash.rs:84-86)
    rng = np.random.default_rng(2)
    a = random_csr(rng, 40, 30, rng.integers(0, 3, size=40) * rng.integers(0, 6, size=40))
    bdeg = rng.integers(0, 8, size=30)
    bdeg[::2] = 0
    b = random_csr(rng, 30, 50, bdeg)
    check_against_oracle(oracle, a, b, gpu_mul(a, b, handle))
    # explicit zeros in the inputs propagate; cancellation zeros are kept
    a = random_csr(rng, 64, 64, 9, dtype=np.int64, zero_frac=0.3)
    b = random_csr(rng, 64, 64, 9, dtype=np.int64, zero_frac=0.3, int_range=1)
    c = gpu_mul(a, b, handle)
    check_against_oracle(oracle, a, b, c)
    assert (c.vals == 0).any()
    # identity
    i = S.CsrMatrix.identity(1000)
    x = random_csr(rng, 1000, 1000, 7)
    c = gpu_mul((1000, 1000, i.offsets, i.indices, i.vals), x, handle)
    assert np.array_equal(c.offsets, x[2]) and np.array_equal(c.indices, x[3]) and np.array_equal(c.vals, x[4])
    # wrapping integers
    big = np.int64(2**62)
    a = (1, 2, [0, 2], [0, 1], np.array([big, big], dtype=np.int64))
    b = (2, 1, [0, 1, 2], [0, 0], np.array([2, 2], dtype=np.int64))
    c = gpu_mul(a, b, handle)
    assert c.vals.tolist() == [0] and c.indices.tolist() == [0]
    # a * b operator (impl Mul for &CsrMatrix): Output = CsrMatrix<T, false>
    A, B = as_csr_matrix(x), as_csr_matrix(x)
    prod = A * B
    assert prod.is_sorted is False and prod.invariants()
    check_against_oracle(oracle, x, x, prod)
    # A * A with the same object: uploaded once
    check_against_oracle(oracle, x, x, A.mul_hash(A, True, handle=handle))


def test_error_behaviour(handle):
    a = S.CsrMatrix.identity(4)
    b = S.CsrMatrix.identity(5)
    with pytest.raises(S.DimensionMismatch):       # the reference panics out-of-bounds (mul_hash.rs:46)
        a.mul_hash(b, handle=handle)
    bad = S.CsrMatrix(2, 2, [1.0, 1.0], [0, 1], [0, 1, 2])
    bad.indices[1] = 9                             # column index >= rows(B): IndexError, not a device fault
    with pytest.raises(IndexError):
        bad.mul_hash(S.CsrMatrix.identity(2), handle=handle)
    L = handle.L
    assert L.spam_spgemm_numeric(handle.h, None, None, 1) == 6      # SPAM_ESTATE: numeric without symbolic
    assert L.spam_dok_to_csr_fetch(handle.h, None, None) == 6
    f32 = S.CsrMatrix.identity(4, dtype=np.float32)
    with pytest.raises(TypeError):
        a.mul_hash(f32, handle=handle)
    # the handle still works after errors
    c = a.mul_hash(a, handle=handle)
    assert c.nnz() == 4
    with pytest.raises(IndexError):
        S.CsrMatrix.from_triplets(2, 2, [0, 2], [0, 0], np.array([1.0, 1.0]), handle=handle)


# ---------------------------------------------------------------------------------------------
# the named configurations
# ---------------------------------------------------------------------------------------------
def test_config1_uniform_10k(oracle, handle):
    a = G.uniform_random(10_000, 10_000, 10, seed=1)
    c = gpu_mul(a, a, handle)
    off, idx, val = check_against_oracle(oracle, a, a, c)
    assert 990_000 < len(idx) < 1_000_000


def test_config2_poisson_full_size(oracle, handle):
    """C2 at BASELINE.json's full size, device-resident path: counts from SURVEY §8, full comparison
    against the oracle (it finishes in seconds), and the row-sum identity (A*A)*1 == A*(A*1)."""
    p = G.poisson2d(2048)
    A = as_csr_matrix(p)
    dA = S.DeviceCsr.upload(A, handle)
    handle.set_timing(True)
    dC = dA.matmul(dA)
    st = handle.stats()
    handle.set_timing(False)
    assert st["flops"] == 104_783_880 and st["nnz_c"] == 54_484_996
    assert st["sym_bin_rows"][MERGE] == 4_194_304 and st["num_bin_rows"][MERGE] == 4_194_304   # sorted B: merge bin
    c = dC.download()
    check_against_oracle(oracle, p, p, c, exact_values=True)
    ones = np.ones(p[0])
    y1 = c.spmv(ones, handle=handle)
    y2 = A.spmv(A.spmv(ones, handle=handle), handle=handle)
    assert np.array_equal(y1, y2)        # small integers: exact in f64
    dC.free()
    dA.free()
    # the same matrix as CsrMatrix<T, false>: every row's entries shuffled.  B's sorted copy is made on the device by two
    # transposes (bucket path with 8192 column buckets, 21 M entries) and the product must be the same, bit for bit.
    rng = np.random.default_rng(7)
    lens = np.diff(p[2].astype(np.int64))
    perm = np.argsort(np.repeat(np.arange(p[0]), lens) + rng.random(len(p[3])), kind="stable")
    dU = S.DeviceCsr.upload(S.CsrMatrix(p[0], p[1], p[4][perm], p[3][perm], p[2], is_sorted=False), handle)
    dT = dU.transpose()
    assert handle.stats()["fallbacks"][4] == 3                        # bucket path
    dT.free()
    dCu = dU.matmul(dU)
    assert handle.stats()["num_bin_rows"][MERGE] == 4_194_304          # multiplied by the sorted copy: merge bin
    cu = dCu.download()
    assert np.array_equal(cu.offsets, c.offsets) and np.array_equal(cu.indices, c.indices)
    assert np.array_equal(cu.vals.view(np.uint8), c.vals.view(np.uint8))
    dCu.free()
    dU.free()


def test_config3_stencil27_reduced(oracle, handle):
    s = G.stencil27(40)
    c = gpu_mul(s, s, handle)
    off, idx, val = check_against_oracle(oracle, s, s, c)
    assert len(idx) == (5 * 40 - 6) ** 3


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_merge_tree_bins(oracle, handle_msort, dtype):
    """SPAM_ESC=3: rows with nnz > 256 and 2 nnz >= products are merged, not hashed (msort.cuh, numeric bins 11..14):
    len(A row) sorted runs, log2 passes of pairwise merge-path merges in shared memory, equal columns folded in the
    reference's product order — structure bit-exact and VALUES bit-exact for every dtype, floats included."""
    h = handle_msort
    rng = np.random.default_rng(1234)
    inner, n = 3000, 60000
    # (1) every product a new column (mostly), A rows of 7 .. 200 entries, B rows of 0 .. 80 entries (empty runs)
    # (rows past 8192 products would take the hash bins, whose float sums are not bit-identical: none here)
    deg = np.array([7, 13, 20, 26, 33, 40, 51, 64, 65, 90, 100, 128, 129, 150, 160, 170] + [3] * 10)
    a = random_csr(rng, len(deg), inner, deg, dtype=dtype, sorted_rows=False)
    b = random_csr(rng, inner, n, rng.integers(0, 81, size=inner), dtype=dtype, sorted_rows=False)
    assert G.spgemm_counts(a, b)[1].max() <= 8192
    c = gpu_mul(a, b, h)
    st = h.stats()
    assert all(x > 0 for x in st["num_bin_rows"][11:15]) and sum(st["num_bin_rows"][4:10]) == 0, st["num_bin_rows"]
    check_against_oracle(oracle, a, b, c, exact_values=True)
    # (2) a third of the columns produced twice: A references row k of B and, for the first third of its entries,
    #     also row k + inner of a B that repeats itself — those entries of C are sums of two products
    deg2 = np.array([12, 21, 42, 60, 90, 120] + [2] * 5)
    a1 = random_csr(rng, len(deg2), inner, deg2, dtype=dtype, sorted_rows=False)
    parts = []
    for lo, hi in zip(a1[2][:-1], a1[2][1:]):
        rowi = a1[3][int(lo):int(hi)]
        parts.append(np.concatenate([rowi, rowi[:len(rowi) // 3] + np.uint64(inner)]))
    idx2 = np.concatenate(parts)
    off2 = np.concatenate([[0], np.cumsum([len(x) for x in parts])]).astype(np.uint64)
    if np.dtype(dtype).kind == "f":
        v2 = rng.uniform(-1, 1, size=len(idx2)).astype(dtype)
    else:
        v2 = rng.integers(-50, 51, size=len(idx2)).astype(dtype)
    a2 = (len(deg2), 2 * inner, off2, idx2, v2)
    b1 = random_csr(rng, inner, n, 45, dtype=dtype, sorted_rows=True)
    b2 = (2 * inner, n, np.concatenate([b1[2], b1[2][1:] + b1[2][-1]]), np.concatenate([b1[3], b1[3]]),
          np.concatenate([b1[4], b1[4][::-1].copy()]))
    c = gpu_mul(a2, b2, h)
    st = h.stats()
    assert sum(st["num_bin_rows"][11:15]) >= 4, st["num_bin_rows"]
    off, idx, val = check_against_oracle(oracle, a2, b2, c, exact_values=True)
    assert len(idx) < G.spgemm_counts(a2, b2)[0] * 0.8
    # (3) B rows of a single entry: as many runs as products (13 passes for 8192 products)
    a3 = random_csr(rng, 6, 40000, np.array([300, 1000, 2000, 4000, 8000, 5]), dtype=dtype, sorted_rows=False)
    b3 = random_csr(rng, 40000, 1 << 22, 1, dtype=dtype)
    c = gpu_mul(a3, b3, h)
    st = h.stats()
    assert all(x > 0 for x in st["num_bin_rows"][11:15]), st["num_bin_rows"]
    check_against_oracle(oracle, a3, b3, c, exact_values=True)
    # (4) power-law columns
    if np.dtype(dtype) == np.float64:
        r = G.rmat(15, 16)
        c = gpu_mul(r, r, h)
        st = h.stats()
        assert all(x > 0 for x in st["num_bin_rows"][11:15]), st["num_bin_rows"]
        check_against_oracle(oracle, r, r, c)     # hash bins also run here: tolerance


def test_config4_rmat_reduced(oracle, handle, handle_esc):
    r = G.rmat(16, 16)
    handle.set_timing(True)
    c = gpu_mul(r, r, handle)
    st = handle.stats()
    handle.set_timing(False)
    check_against_oracle(oracle, r, r, c)
    # the bucket-sort bins on power-law columns (crowded level-1 buckets are split a second time)
    c = gpu_mul(r, r, handle_esc)
    se = handle_esc.stats()
    check_against_oracle(oracle, r, r, c)
    assert sum(se["num_bin_rows"][11:16]) > 0 and se["num_bin_rows"][15] > 0, se["num_bin_rows"]
    assert st["sym_bin_rows"][HEAVY] > 0, st["sym_bin_rows"]      # power-law rows reach the global-table bin
    assert sum(st["num_bin_rows"][11:16]) + st["num_bin_rows"][HEAVY] > 0, st["num_bin_rows"]


# ---------------------------------------------------------------------------------------------
# the rarely taken code paths: each is forced by a hand-built shape and asserted through spam_stats.fallbacks
# ---------------------------------------------------------------------------------------------
def _csr_from_rows(rows_cols, ncols, rng, dtype=np.float64):
    off = np.zeros(len(rows_cols) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in rows_cols])
    idx = np.concatenate([np.asarray(c, np.uint64) for c in rows_cols]) if rows_cols else np.empty(0, np.uint64)
    if np.dtype(dtype).kind == "f":
        val = rng.uniform(-1, 1, size=len(idx)).astype(dtype)
        val[val == 0] = 0.5
    else:
        val = rng.integers(1, 50, size=len(idx)).astype(dtype)
    return len(rows_cols), ncols, off, idx, val


@pytest.mark.parametrize("sorted_b", [False, True])
def test_fallback_wide_columns_warp_bitonic(oracle, handle, sorted_b):
    """The fuzz target's shape space (fuzz/fuzz_targets/mul_hash.rs:15-19: n up to 2^31): with cols(B) = 2^31 - 2 a
    column no longer packs with its index into 32 bits, so the one-warp bins sort with the shared-memory bitonic
    network (rowhash.cuh warp_bitonic_sort)."""
    rng = np.random.default_rng(2024)
    n = 2**31 - 2
    inner = 200
    brows = [np.unique(rng.integers(0, n, size=int(k))) for k in rng.integers(0, 25, size=inner)]
    brows[7] = np.array([0, 1, n - 2, n - 1])                  # both ends of the column range
    if not sorted_b:
        brows = [rng.permutation(c) for c in brows]
    b = _csr_from_rows(brows, n, rng)
    a = random_csr(rng, 300, inner, rng.integers(0, 22, size=300), sorted_rows=False)
    c = gpu_mul(a, b, handle)
    st = handle.stats()
    check_against_oracle(oracle, a, b, c)
    assert sum(st["num_bin_rows"][1:4]) > 0 and st["fallbacks"][0] > 0, (st["num_bin_rows"], st["fallbacks"])


@pytest.mark.parametrize("dtype", [np.float64, np.int64])
def test_fallback_team_bucket_overflow(oracle, handle, dtype):
    """A team-bin row whose columns are 600 consecutive ids plus one far away: the order-preserving buckets over the
    row's column range put all 600 into one bucket (> 512), so the drain compacts in shared memory and runs the
    block-wide bitonic network (rowhash.cuh, NW > 1 fallback).  Three copies of every product keep the row on
    the hash path (it compresses 3x)."""
    rng = np.random.default_rng(7)
    far = 2**30
    base = [np.arange(0, 300), np.arange(300, 600), np.array([far])]
    b = _csr_from_rows(base * 3 + [np.arange(5, 9)], far + 1, rng, dtype)
    a = _csr_from_rows([np.arange(9), np.array([9]), np.arange(3)], 10, rng, dtype)
    c = gpu_mul(a, b, handle)
    st = handle.stats()
    check_against_oracle(oracle, a, b, c)
    assert st["num_bin_rows"][5] >= 1 and st["fallbacks"][1] >= 1, (st["num_bin_rows"], st["fallbacks"])


@pytest.mark.parametrize("dtype", [np.float64, np.int32])
def test_fallback_heavy_bucket_overflow(oracle, handle, dtype):
    """A global-table row (nnz > 8192) with 3000 consecutive columns in one drain bucket (> 2048): in-place
    compaction and the global-memory bitonic network of k_num_heavy.  Every column is produced three times."""
    rng = np.random.default_rng(8)
    ncols = 2**30
    spread = np.unique(rng.integers(1 << 20, ncols, size=6500))
    cols = np.concatenate([np.arange(70000, 73000), spread])     # 3000 consecutive ids: one 65536-wide bucket
    chunks = np.array_split(cols, 40)
    b = _csr_from_rows(chunks * 3, ncols, rng, dtype)
    a = _csr_from_rows([np.arange(120), np.arange(40)], 120, rng, dtype)
    c = gpu_mul(a, b, handle)
    st = handle.stats()
    check_against_oracle(oracle, a, b, c)
    assert len(cols) > 8192 and st["num_bin_rows"][HEAVY] >= 1 and st["fallbacks"][2] >= 1, (st["num_bin_rows"], st["fallbacks"])


def test_config5_rectangular_i64_and_dok(oracle, handle):
    a = G.uniform_random(100_000, 400_000, 8, seed=5, dtype=np.int64, int_range=1 << 15)
    at = G.transpose(a)
    c = gpu_mul(a, at, handle)
    check_against_oracle(oracle, a, at, c)
    tr, tc, tv = G.triplets_with_rewrites(a, seed=5)
    got = S.CsrMatrix.from_triplets(a[0], a[1], tr, tc, tv, handle=handle)
    off, idx, val = oracle.dok_to_csr(a[0], a[1], tr, tc, tv)
    assert got.invariants()
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)


# ---------------------------------------------------------------------------------------------
# SpMV, DOK -> CSR, device-resident helpers
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_spmv(oracle, handle, dtype):
    rng = np.random.default_rng(8)
    for rows, cols, deg in [(1, 1, 1), (500, 300, 3), (2000, 2500, 40), (64, 5000, 900)]:
        a = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, sorted_rows=False)
        if np.dtype(dtype).kind == "f":
            x = rng.uniform(-1, 1, size=cols).astype(dtype)
        else:
            x = rng.integers(-50, 50, size=cols).astype(dtype)
        y = as_csr_matrix(a, False).spmv(x, handle=handle)
        want = oracle.spmv(rows, cols, a[2], a[3], a[4], x)
        if np.dtype(dtype).kind == "f":
            sabs = oracle.spmv(rows, cols, a[2], a[3], np.abs(a[4]), np.abs(x)).astype(np.float64)
            assert np.all(np.abs(y.astype(np.float64) - want.astype(np.float64)) <= TOL[np.dtype(dtype)] * sabs)
        else:
            assert np.array_equal(y, want)
    with pytest.raises(S.DimensionMismatch):
        S.CsrMatrix.identity(3).spmv(np.ones(4), handle=handle)


def test_spmv_tma_kernel(oracle):
    """SPAM_SPMV_TMA=1: the persistent kernel that fetches col_idx / values with cp.async.bulk two stages deep.  Row
    blocks with more than one chunk of 1536 entries, empty row blocks, nnz not a multiple of 4, more row blocks than
    resident thread blocks (> 148 * 6 * 256 rows); f64 sums bit-identical to the oracle (same order, unfused)."""
    os.environ["SPAM_SPMV_TMA"] = "1"
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_SPMV_TMA"]
    try:
        rng = np.random.default_rng(88)
        cases = []
        for dtype in ALL_DTYPES:
            for rows, cols, deg in [(1, 1, 1), (700, 300, 3), (3000, 2500, 40), (1000, 5000, 64)]:
                cases.append(random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype))
        deg = rng.integers(0, 4, size=300_000)
        deg[1000:3000] = 0                                  # whole row blocks without entries
        cases.append(random_csr(rng, 300_000, 4000, deg, dtype=np.float64))
        cases.append(G.poisson2d(300, dtype=np.float64))
        for a in cases:
            dtype = a[4].dtype
            x = (rng.uniform(-1, 1, size=a[1]) if dtype.kind == "f" else rng.integers(-50, 50, size=a[1])).astype(dtype)
            y = as_csr_matrix(a, True).spmv(x, handle=h)
            want = oracle.spmv(a[0], a[1], a[2], a[3], a[4], x)
            assert np.array_equal(y, want), (a[0], a[1], dtype)
    finally:
        h.close()


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_dok_to_csr(oracle, handle, dtype):
    rng = np.random.default_rng(21)
    for rows, cols, n in [(1, 1, 0), (1, 1, 3), (5, 5, 40), (300, 70_000, 20_000), (70_000, 3, 50_000)]:
        ri = rng.integers(0, rows, n)
        ci = rng.integers(0, cols, n)
        v = rng.integers(-2, 3, n).astype(dtype)          # many zeros: deletes and re-inserts
        got = S.CsrMatrix.from_triplets(rows, cols, ri, ci, v, handle=handle)
        off, idx, val = oracle.dok_to_csr(rows, cols, ri, ci, v)
        assert got.invariants()
        assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx)
        assert np.array_equal(got.vals, val)
    d = S.DokMatrix.new((4, 4), dtype)
    d.set_element((3, 1), 2)
    d.set_element((0, 0), 1)
    d.set_element((3, 1), 0)
    d.set_element((2, 2), 5)
    m = S.CsrMatrix.from_dok(d, handle=handle)
    assert m.offsets.tolist() == [0, 1, 1, 2, 2] and m.indices.tolist() == [0, 2] and m.to_dok().entries == d.entries


def test_dok_and_transpose_take_both_paths(oracle, handle_nobucket):
    """Without the bucket path (SPAM_DOK_BUCKET=0; it is what the bucket path falls back to): DOK -> CSR and transpose
    pick the counting path when no row (column) holds more than 32 entries and the stable
    radix sort otherwise (spam_stats.fallbacks[4] = 1 / 2); both are bit-exact, including a stream that rewrites one
    key many times (last write wins, a final zero deletes)."""
    rng = np.random.default_rng(77)
    rows, cols = 5000, 9000
    # short rows with rewrites and deletions: counting path
    a = random_csr(rng, rows, cols, rng.integers(0, 13, size=rows), dtype=np.int64)
    tr, tc, tv = G.triplets_with_rewrites(a, seed=9, dup_frac=0.3, zero_frac=0.1)
    got = S.CsrMatrix.from_triplets(rows, cols, tr, tc, tv, handle=handle_nobucket)
    assert handle_nobucket.stats()["fallbacks"][4] == 1
    off, idx, val = oracle.dok_to_csr(rows, cols, tr, tc, tv)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    # exactly 32 triplets in one row, all on two keys: still the counting path
    tr2 = np.concatenate([tr, np.full(32, rows - 1)]); tc2 = np.concatenate([tc, np.tile([5, 6], 16)])
    keep = tr2 != rows - 1
    keep[-32:] = True
    tr2, tc2 = tr2[keep], tc2[keep]
    tv2 = np.concatenate([tv[keep[:len(tv)]], np.arange(1, 33)])
    tv2[-1] = 0                                       # the last write of key (rows-1, 6) is a zero: deleted
    got = S.CsrMatrix.from_triplets(rows, cols, tr2, tc2, tv2, handle=handle_nobucket)
    assert handle_nobucket.stats()["fallbacks"][4] == 1
    off, idx, val = oracle.dok_to_csr(rows, cols, tr2, tc2, tv2)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    assert got.get_element((rows - 1, 5)) == 31 and got.get_element((rows - 1, 6)) is None
    # one more triplet in that row (33): radix path, same answer as the oracle
    tr3, tc3, tv3 = np.append(tr2, rows - 1), np.append(tc2, 7), np.append(tv2, 9)
    got = S.CsrMatrix.from_triplets(rows, cols, tr3, tc3, tv3, handle=handle_nobucket)
    assert handle_nobucket.stats()["fallbacks"][4] == 2
    off, idx, val = oracle.dok_to_csr(rows, cols, tr3, tc3, tv3)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    # transpose: columns of at most 32 entries -> counting path; a dense column -> radix path
    A = as_csr_matrix(a)
    t = A.transpose(handle=handle_nobucket)
    longest = int(np.bincount(a[3].astype(np.int64), minlength=cols).max())
    assert longest <= 32 and handle_nobucket.stats()["fallbacks"][4] == 1
    want = oracle.transpose(a)
    assert np.array_equal(t.offsets, want[0]) and np.array_equal(t.indices, want[1]) and np.array_equal(t.vals, want[2])
    dense = random_csr(rng, 400, 50, np.full(400, 20), dtype=np.float64, sorted_rows=False)   # 160 per column
    t = as_csr_matrix(dense, is_sorted=False).transpose(handle=handle_nobucket)
    assert handle_nobucket.stats()["fallbacks"][4] == 2
    want = oracle.transpose(dense)
    assert np.array_equal(t.offsets, want[0]) and np.array_equal(t.indices, want[1]) and np.array_equal(t.vals, want[2])


def test_dok_and_transpose_bucket_path(oracle, handle):
    """The default path (bucket.cuh, spam_stats.fallbacks[4] = 3): one partition pass into buckets of consecutive rows
    (columns), one build pass in shared memory.  Bit-exact against the oracle for every dtype, for streams with
    rewrites and deletions, rows from empty to several hundred triplets, a row count that is not a multiple of the
    bucket width, and wide column spaces; a row past 512 triplets or a crowded bucket hands the build to the
    counting / radix paths (fallbacks[4] = 1 / 2) with the same answer."""
    rng = np.random.default_rng(404)
    shapes = [(5000, 9000, 13), (1, 70000, 300), (70001, 3, 3), (3000, (1 << 31) - 2, 9), (1 << 17, 1 << 22, 6),
              (40000, 1 << 26, 20)]
    for k, (rows, cols, per) in enumerate(shapes):
        dtype = [np.int64, np.float64, np.float32, np.int32][k % 4]
        lens = rng.integers(0, min(per, cols) + 1, size=rows)
        a = random_csr(rng, rows, cols, lens, dtype=dtype)
        tr, tc, tv = G.triplets_with_rewrites(a, seed=k, dup_frac=0.3, zero_frac=0.1)
        got = S.CsrMatrix.from_triplets(rows, cols, tr, tc, tv, handle=handle)
        assert handle.stats()["fallbacks"][4] == 3, (rows, cols, handle.stats()["fallbacks"])
        off, idx, val = oracle.dok_to_csr(rows, cols, tr, tc, tv)
        assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx)
        assert np.array_equal(got.vals.view(np.uint8), val.view(np.uint8)), (rows, cols)
        if cols <= (1 << 22):   # the transpose keeps cols + 1 offsets
            t = as_csr_matrix(a).transpose(handle=handle)
            want = oracle.transpose(a)
            assert np.array_equal(t.offsets, want[0]) and np.array_equal(t.indices, want[1])
            assert np.array_equal(t.vals.view(np.uint8), want[2].view(np.uint8))
    # one key rewritten 400 times (last write wins), another ending in a zero (deleted): still the bucket path
    rows, cols = 5000, 9000
    a = random_csr(rng, rows, cols, rng.integers(0, 13, size=rows), dtype=np.int64)
    tr, tc, tv = G.triplets_with_rewrites(a, seed=9, dup_frac=0.3, zero_frac=0.1)
    keep = tr != rows - 1
    tr2 = np.concatenate([tr[keep], np.full(400, rows - 1)])
    tc2 = np.concatenate([tc[keep], np.tile([5, 6], 200)])
    tv2 = np.concatenate([tv[keep], np.arange(1, 401)])
    tv2[-1] = 0
    got = S.CsrMatrix.from_triplets(rows, cols, tr2, tc2, tv2, handle=handle)
    assert handle.stats()["fallbacks"][4] == 3
    off, idx, val = oracle.dok_to_csr(rows, cols, tr2, tc2, tv2)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    assert got.get_element((rows - 1, 5)) == 399 and got.get_element((rows - 1, 6)) is None
    # 600 triplets in one row: past the all-pairs limit -> flag -> the radix path, same answer
    tr3 = np.concatenate([tr2, np.full(200, rows - 1)]); tc3 = np.concatenate([tc2, np.arange(100, 300)])
    tv3 = np.concatenate([tv2, np.arange(1, 201)])
    got = S.CsrMatrix.from_triplets(rows, cols, tr3, tc3, tv3, handle=handle)
    assert handle.stats()["fallbacks"][4] == 2
    off, idx, val = oracle.dok_to_csr(rows, cols, tr3, tc3, tv3)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    # every triplet in the first 64 of 200000 rows: the buckets of those rows overflow -> counting path
    rows, cols = 200000, 1000
    tr = rng.integers(0, 64, size=300000).astype(np.uint64); tc = rng.integers(0, cols, size=300000).astype(np.uint64)
    tv = rng.integers(1, 100, size=300000).astype(np.int64)
    got = S.CsrMatrix.from_triplets(rows, cols, tr, tc, tv, handle=handle)
    assert handle.stats()["fallbacks"][4] in (1, 2)
    off, idx, val = oracle.dok_to_csr(rows, cols, tr, tc, tv)
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    # a dense column block (160 entries per column) and an out-of-range column
    dense = random_csr(rng, 400, 50, np.full(400, 20), dtype=np.float64, sorted_rows=False)
    t = as_csr_matrix(dense, is_sorted=False).transpose(handle=handle)
    assert handle.stats()["fallbacks"][4] == 3
    want = oracle.transpose(dense)
    assert np.array_equal(t.offsets, want[0]) and np.array_equal(t.indices, want[1]) and np.array_equal(t.vals, want[2])
    with pytest.raises(Exception):
        S.CsrMatrix.from_triplets(10, 10, np.array([1, 10], np.uint64), np.array([1, 1], np.uint64),
                                  np.array([1, 2], np.int64), handle=handle)


def test_rows_to_parts_and_row_slices(oracle, handle):
    """Multi-GPU plumbing on one GPU: the flop-balanced partition equals rows_to_threads
    (mul_hash.rs:51-62) and the row-block products concatenate to the full product."""
    r = G.rmat(13, 12)
    A = as_csr_matrix(r)
    dA = S.DeviceCsr.upload(A, handle)
    full = dA.matmul(dA).download()
    _, _, sabs = oracle.mul_hash(r[:4] + (np.abs(r[4]),), r[:4] + (np.abs(r[4]),), True)
    for parts in (1, 2, 3, 8):
        starts, total = dA.rows_to_parts(dA, parts)
        flop, ro = oracle.rows_to_threads(r[0], r[2], r[3], r[2], parts)
        assert np.array_equal(starts, ro) and total == int(flop.sum())
        # the device-time-balanced variant: same formula on cost_i = flop_i * w(flop_i)
        cstarts, ctotal = dA.rows_to_parts(dA, parts, balance="cost")
        assert np.array_equal(cstarts, D.partition_rows_from_flops(D.row_cost(flop), parts)) and ctotal == total
        offs, idxs, vals = [np.zeros(1, np.uint64)], [], []
        for t in range(parts):
            blk = dA.slice_rows(int(starts[t]), int(starts[t + 1]))
            c = blk.matmul(dA).download()
            offs.append(c.offsets[1:] + offs[-1][-1])     # offset-fixed row_ptr shard
            idxs.append(c.indices)
            vals.append(c.vals)
            blk.free()
        assert np.array_equal(np.concatenate(offs), full.offsets)
        assert np.array_equal(np.concatenate(idxs), full.indices)
        # block products accumulate in another (atomic) order: same tolerance as against the oracle
        assert np.all(np.abs(np.concatenate(vals) - full.vals) <= 2 * TOL[np.dtype(np.float64)] * sabs)
    dA.free()


def _check_unsorted(oracle, a, b, c, exact=False):
    off, idx, val = oracle.mul_hash(a, b, False)          # B2 = false: the reference's slot order
    assert np.array_equal(c.offsets, off), "row_ptr differs"
    assert np.array_equal(c.indices, idx), "col_idx (slot order) differs"
    dt = np.dtype(val.dtype)
    if dt.kind != "f" or exact:
        assert np.array_equal(c.vals, val)
    else:
        sabs = oracle.mul_hash(a[:4] + (np.abs(a[4]),), b[:4] + (np.abs(b[4]),), False)[2]
        assert np.all(np.abs(c.vals.astype(np.float64) - val.astype(np.float64)) <= TOL[dt] * sabs.astype(np.float64))


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_unsorted_output_in_reference_slot_order(oracle, handle, dtype):
    """B2 = false (what `&a * &b` and the reference's bench return, mul_hash.rs:176-186): col_idx must come back in
    the slot order of linprobe's map under the reference's insertion order, bit for bit — rows short enough for the
    thread-per-row replay (<= 32 columns), block-per-row priority insertion (<= 4096) and the global-scratch kernel."""
    g = GOLD["map_slot_order"]
    a = (1, 4, g["a"]["offsets"], g["a"]["indices"], np.array(g["a"]["vals"], dtype=dtype))
    b = (4, 4, g["b"]["offsets"], g["b"]["indices"], np.array(g["b"]["vals"], dtype=dtype))
    c = as_csr_matrix(a, False).mul_hash(as_csr_matrix(b, False), sorted_output=False, handle=handle, reference_order=True)
    assert c.indices.tolist() == g["unsorted"]["indices"]
    rng = np.random.default_rng(2718)
    # many short rows with colliding columns (multiples of the table size), unsorted inputs
    a = random_csr(rng, 400, 300, rng.integers(0, 7, size=400), dtype=dtype, sorted_rows=False)
    b = random_csr(rng, 300, 4096, rng.integers(0, 9, size=300), dtype=dtype, sorted_rows=False)
    b[3][:] = (b[3] // 16) * 16 % 4096 + (b[3] % 2)          # columns cluster on few hash slots
    b = _dedupe_rows(b)
    c = as_csr_matrix(a, False).mul_hash(as_csr_matrix(b, False), sorted_output=False, handle=handle, reference_order=True)
    _check_unsorted(oracle, a, b, c)
    # medium and long rows: up to ~20 000 columns in a row, compressing and not
    deg = np.array([0, 3, 12, 40, 90, 200, 520, 130, 7] + [25] * 30)
    a = random_csr(rng, len(deg), 600, deg, dtype=dtype, sorted_rows=False)
    b = random_csr(rng, 600, 30000, 40, dtype=dtype, sorted_rows=False)
    A, B = as_csr_matrix(a, False), as_csr_matrix(b, False)
    c = A.mul_hash(B, sorted_output=False, handle=handle, reference_order=True)
    _check_unsorted(oracle, a, b, c)
    assert max(np.diff(c.offsets.astype(np.int64))) > 4096
    assert c.invariants()
    # sorted inputs (merge bin) and the device-resident entry point
    p = G.poisson2d(40, dtype=dtype) if np.dtype(dtype).kind == "f" else None
    if p is not None:
        dA = S.DeviceCsr.upload(as_csr_matrix(p), handle)
        dC = dA.matmul(dA, reference_order=True)
        _check_unsorted(oracle, p, p, dC.download(is_sorted=False), exact=True)
        dC.free(); dA.free()


def test_fuzz_target_shape_space(oracle, handle):
    """The reference's fuzz target (fuzz/fuzz_targets/mul_hash.rs:11-50): f64, l and m in [1, 256], n in [1, 2^31],
    at most 1000 random set_element calls per matrix (spam_matrix/src/arbitrary.rs:13-19); it asserts invariants()
    always.  Here: 60 seeded draws from that space (n log-uniform), both output orders against the oracle."""
    rng = np.random.default_rng(20260)
    for case in range(60):
        l, m = int(rng.integers(1, 257)), int(rng.integers(1, 257))
        n = int(min(2 ** 31 - 2, max(1, round(2 ** rng.uniform(0, 31)))))
        mats = []
        for rows, cols in ((l, m), (m, n)):
            d = S.DokMatrix.new((rows, cols))
            for _ in range(int(rng.integers(0, 1001))):
                d.set_element((int(rng.integers(0, rows)), int(rng.integers(0, cols))), float(rng.normal()))
            keys = list(d.entries)                      # insertion order: rows come out unsorted
            order = np.argsort([k[0] for k in keys], kind="stable")
            rr = np.array([keys[i][0] for i in order], dtype=np.int64)
            off = np.zeros(rows + 1, np.uint64)
            np.cumsum(np.bincount(rr, minlength=rows), out=off[1:])
            mats.append((rows, cols, off, np.array([keys[i][1] for i in order], np.uint64),
                         np.array([d.entries[keys[i]] for i in order], np.float64)))
        a, b = mats
        A, B = as_csr_matrix(a, False), as_csr_matrix(b, False)
        c = A.mul_hash(B, sorted_output=True, handle=handle)
        check_against_oracle(oracle, a, b, c)
        c = A.mul_hash(B, sorted_output=False, handle=handle, reference_order=True)
        assert c.invariants()
        _check_unsorted(oracle, a, b, c)


def _dedupe_rows(m):
    rows, cols, off, idx, val = m
    noff, nidx, nval = [0], [], []
    for r in range(rows):
        lo, hi = int(off[r]), int(off[r + 1])
        _, first = np.unique(idx[lo:hi], return_index=True)
        first = np.sort(first)                        # keep the original (unsorted) order of the survivors
        nidx.append(idx[lo:hi][first]); nval.append(val[lo:hi][first])
        noff.append(noff[-1] + len(first))
    return rows, cols, np.array(noff, np.uint64), np.concatenate(nidx), np.concatenate(nval)


def test_gathered_product_on_one_rank(oracle):
    """spam_spgemm_gathered with a one-rank communicator: the sub-block pipeline (row views of A, offset-fixed row_ptr,
    numeric kernels writing into the gather buffers) without peers.  The multi-rank exchange itself is checked by
    tests/multi_gpu_check.py under torchrun and by bench.py's gathered_parity at every N > 1."""
    h = S.Handle(0)
    try:
        h.comm_init(S.comm_unique_id(), 0, 1)
        assert h.comm_info() == {"rank": 0, "world": 1, "peer_mapped": True}
        rng = np.random.default_rng(12)
        mats = [G.rmat(13, 12), G.poisson2d(80), random_csr(rng, 700, 700, rng.integers(0, 40, size=700), sorted_rows=False),
                random_csr(rng, 3, 3, 2, dtype=np.int64)]
        for m in mats:
            dA = S.DeviceCsr.upload(as_csr_matrix(m, is_sorted=False), h)
            for nsub, mode in ((1, 0), (3, 0), (16, 1)):
                g = dA.matmul_gathered(dA, 0, m[0], nsub=nsub, mode=mode)
                c = g.download()
                g.free()
                check_against_oracle(oracle, m, m, c)
            # a second, smaller product reuses the buffers; a block that does not tile the matrix is refused
            with pytest.raises(S.SpamError):
                dA.matmul_gathered(dA, 1, m[0], nsub=1)
            dA.free()
        # y = A x through the gathered entry point
        import ctypes as C_
        m = mats[0]
        dA = S.DeviceCsr.upload(as_csr_matrix(m), h)
        x = rng.uniform(-1, 1, size=m[1])
        want = oracle.spmv(m[0], m[1], m[2], m[3], m[4], x)
        dx = S.DeviceCsr.upload(S.CsrMatrix(1, m[1], x, np.arange(m[1]), [0, m[1]]), h)   # a device buffer holding x
        dy = S.DeviceCsr.upload(S.CsrMatrix(1, m[0], np.zeros(m[0]), np.arange(m[0]), [0, m[0]]), h)
        dA.spmv_gathered(dx.info()["d_val"], dy.info()["d_val"], [m[0]])
        y = dy.download().vals
        sabs = oracle.spmv(m[0], m[1], m[2], m[3], np.abs(m[4]), np.abs(x))
        assert np.all(np.abs(y - want) <= 1e-12 * sabs)
        for d in (dA, dx, dy):
            d.free()
        # DOK -> CSR through the sharded entry point (one rank owns every row): partition, self-copy, local build
        u = G.uniform_random(3000, 5000, 7, seed=8, dtype=np.int64, int_range=40)
        tr, tc, tv = G.triplets_with_rewrites(u, seed=3, dup_frac=0.2, zero_frac=0.05)
        n = len(tv)

        def dev_buf(a64):      # a device buffer with these 8-byte words (values of a 1 x n matrix)
            return S.DeviceCsr.upload(S.CsrMatrix(1, max(1, n), a64.view(np.int64), np.arange(n), [0, n]), h)
        br, bc, bv = dev_buf(tr), dev_buf(tc), dev_buf(tv)
        blk, r0 = S.DeviceCsr.from_triplets_sharded(h, np.int64, u[0], u[1], n, br.info()["d_val"], bc.info()["d_val"],
                                                    bv.info()["d_val"])
        got = blk.download()
        off, idx, val = oracle.dok_to_csr(u[0], u[1], tr, tc, tv)
        assert r0 == 0 and np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
        for d in (blk, br, bc, bv):
            d.free()
    finally:
        h.close()


def test_host_path_uploads_only_the_referenced_band_of_b(oracle, handle):
    """A row block of a banded matrix times the whole matrix through the host C ABI: only the rows of B that
    A's columns reference travel over PCIe (the other rows are empty on the device), the product is the same."""
    p = G.poisson2d(64)
    full = as_csr_matrix(p)
    for r0, r1 in ((640, 1280), (0, 64), (4000, 4096), (100, 101)):
        lo, hi = int(p[2][r0]), int(p[2][r1])
        a = (r1 - r0, p[1], (p[2][r0:r1 + 1] - p[2][r0]).astype(np.uint64), p[3][lo:hi].copy(), p[4][lo:hi].copy())
        c = as_csr_matrix(a).mul_hash(full, sorted_output=True, handle=handle)
        check_against_oracle(oracle, a, p, c, exact_values=True)     # merge bin: bit-identical floats
        h2d = handle.stats()["bytes_h2d"]
        whole_b = (p[0] + 1) * 8 + len(p[3]) * 16
        assert h2d < 0.5 * whole_b, (h2d, whole_b)
    # a block that references (almost) every row of B takes the plain upload
    a = random_csr(np.random.default_rng(5), 50, p[0], np.full(50, 40), dtype=np.float64)
    c = as_csr_matrix(a).mul_hash(full, sorted_output=True, handle=handle)
    check_against_oracle(oracle, a, p, c)
    assert handle.stats()["bytes_h2d"] > (p[0] + 1) * 8 + len(p[3]) * 16
    # A with no entries at all, and A whose only column is the last row of B
    e = S.CsrMatrix.new((3, p[0])).mul_hash(full, sorted_output=True, handle=handle)
    assert e.nnz() == 0 and e.offsets.tolist() == [0, 0, 0, 0]
    a = (2, p[0], np.array([0, 1, 1], np.uint64), np.array([p[0] - 1], np.uint64), np.array([2.0]))
    check_against_oracle(oracle, a, p, as_csr_matrix(a).mul_hash(full, sorted_output=True, handle=handle), exact_values=True)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_transpose(oracle, handle, dtype):
    """Matrix::transpose of CsrMatrix (lib.rs:256-264) through the host C ABI and on the device: bit-exact
    against the oracle (values are moved, never recomputed), explicit zeros kept, unsorted input rows,
    rows longer than a warp, empty rows and columns, a single element."""
    rng = np.random.default_rng(21)
    cases = [(1, 1, 1, True), (9, 4, 3, True), (4, 900, 300, False), (700, 6, 5, False), (300, 300, 0, True),
             (2000, 1500, 24, False), (64, 5000, 100, True)]
    for rows, cols, deg, srt in cases:
        a = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, sorted_rows=srt, zero_frac=0.1)
        want = oracle.transpose(a)
        A = as_csr_matrix(a)
        got = A.transpose(handle=handle)
        assert (got.rows(), got.cols()) == (cols, rows) and got.invariants()
        assert np.array_equal(got.offsets, want[0]) and np.array_equal(got.indices, want[1])
        assert np.array_equal(got.vals.view(np.uint8), want[2].view(np.uint8))
        dA = S.DeviceCsr.upload(A, handle)
        dT = dA.transpose()
        dev = dT.download()
        assert np.array_equal(dev.offsets, want[0]) and np.array_equal(dev.indices, want[1])
        assert np.array_equal(dev.vals.view(np.uint8), want[2].view(np.uint8))
        # the transposed matrix is a valid right-hand side: A * A^T against the oracle (C5's shape of product)
        if rows * deg and rows <= 2000:
            c = dA.matmul(dT).download()
            check_against_oracle(oracle, a, (cols, rows) + want, c)
        dT.free(); dA.free()
    bad = S.CsrMatrix(2, 2, np.array([1, 1], dtype=dtype), [0, 1], [0, 1, 2])
    bad.indices[1] = 7
    with pytest.raises(IndexError):
        bad.transpose(handle=handle)
    bad2 = S.CsrMatrix(3, 2, np.array([1, 1], dtype=dtype), [0, 1], [0, 2, 1, 2])   # row_ptr not monotone
    with pytest.raises(Exception):
        bad2.transpose(handle=handle)


@pytest.mark.parametrize("tma", ["1", "0"])
def test_elementwise_span_paths(oracle, tma):
    """k_ewise_fill_tma stages the spans of a block when they fit the capacities taken from the mean row lengths:
    skewed matrices (a few blocks of long rows among short ones: those blocks walk global memory), lengths that are not
    multiples of 4 entries, empty operands; and the same with SPAM_EWISE_TMA=0 (k_ewise_fill)."""
    os.environ["SPAM_EWISE_TMA"] = tma
    try:
        h = S.Handle(0)
    finally:
        del os.environ["SPAM_EWISE_TMA"]
    try:
        rng = np.random.default_rng(97)
        for trial in range(6):
            rows, cols = 5000 + trial, 4000
            da = rng.integers(0, 6, size=rows)
            db = rng.integers(0, 4, size=rows)
            da[1000:1300] = rng.integers(100, 400, size=300)      # spans far beyond the capacities
            db[3000:3100] = 300
            if trial == 4:
                db[:] = 0                                         # B empty
            if trial == 5:
                da[:] = 0
            dtype = (np.float64, np.int64, np.float32, np.int32, np.float64, np.float64)[trial]
            a = random_csr(rng, rows, cols, da, dtype=dtype)
            b = random_csr(rng, rows, cols, db, dtype=dtype)
            for op in ("add", "sub"):
                want = oracle.ewise(a, b, op, True)
                dA, dB = S.DeviceCsr.upload(as_csr_matrix(a), h), S.DeviceCsr.upload(as_csr_matrix(b), h)
                dC = dA.add(dB) if op == "add" else dA.sub(dB)
                got = dC.download()
                assert np.array_equal(got.offsets, want[0]) and np.array_equal(got.indices, want[1])
                assert np.array_equal(got.vals.view(np.uint8), want[2].view(np.uint8))
                for d in (dA, dB, dC):
                    d.free()
    finally:
        h.close()


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("op", ["add", "sub"])
def test_elementwise_add_sub(oracle, handle, dtype, op):
    """impl Add / Sub for CsrMatrix (apply_elementwise, lib.rs:83-149) through the host C ABI and on the device:
    bit-exact against the oracle (one rounding per entry, integers wrap), the union pattern with cancellation
    and explicit zeros kept, unsorted operands (put in order by two transposes), A op A, shape mismatch."""
    rng = np.random.default_rng(31)
    info = np.iinfo(dtype) if np.dtype(dtype).kind == "i" else None
    for rows, cols, deg, srt in ((1, 1, 1, True), (40, 30, 6, True), (3000, 2500, 20, True), (500, 700, 90, False),
                                 (64, 64, 0, True)):
        a = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, sorted_rows=srt, zero_frac=0.1,
                       int_range=(info.max // 2 + 7) if info else 50)
        b = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, sorted_rows=srt, zero_frac=0.1,
                       int_range=(info.max // 2 + 7) if info else 50)
        want = oracle.ewise(a, b, op, srt)
        A, B = as_csr_matrix(a, is_sorted=srt), as_csr_matrix(b, is_sorted=srt)
        got = A.add(B, handle=handle) if op == "add" else A.sub(B, handle=handle)
        assert (got.rows(), got.cols()) == (rows, cols)
        assert np.array_equal(got.offsets, want[0]) and np.array_equal(got.indices, want[1])
        assert np.array_equal(got.vals.view(np.uint8), want[2].view(np.uint8))
        dA, dB = S.DeviceCsr.upload(A, handle), S.DeviceCsr.upload(B, handle)
        dC = dA.add(dB) if op == "add" else dA.sub(dB)
        dev = dC.download()
        # the device entry point applies the IS_SORTED = true rule to the operands' rows put in column order
        sa = a if srt else (rows, cols) + oracle.transpose((cols, rows) + oracle.transpose(a))
        sb = b if srt else (rows, cols) + oracle.transpose((cols, rows) + oracle.transpose(b))
        wsorted = oracle.ewise(sa, sb, op, True)
        assert np.array_equal(dev.offsets, wsorted[0]) and np.array_equal(dev.indices, wsorted[1])
        assert np.array_equal(dev.vals, wsorted[2])
        # A op A: add doubles, sub leaves the pattern full of cancellation zeros
        dS = dA.add(dA) if op == "add" else dA.sub(dA)
        same = dS.download()
        wself = oracle.ewise(sa, sa, op, True)
        assert np.array_equal(same.offsets, wself[0]) and np.array_equal(same.indices, wself[1])
        assert np.array_equal(same.vals, wself[2]) and same.nnz() == len(a[3])
        for d in (dA, dB, dC, dS):
            d.free()
    # left-only -0.0 under add: +0.0 with IS_SORTED = true, kept with IS_SORTED = false (lib.rs:114 vs 119-137)
    if np.dtype(dtype).kind == "f":
        a = S.CsrMatrix(1, 2, np.array([-0.0], dtype), [0], [0, 1])
        b = S.CsrMatrix(1, 2, np.array([2.0], dtype), [1], [0, 1])
        assert not np.signbit(a.add(b, handle=handle).vals[0])
        a.is_sorted = False
        assert np.signbit(a.add(b, handle=handle).vals[0])
    with pytest.raises(S.DimensionMismatch):
        S.CsrMatrix.identity(3, dtype=dtype).add(S.CsrMatrix.identity(4, dtype=dtype), handle=handle)


def test_scan_sizes_through_dok_row_ptr(oracle, handle):
    """The look-back scan at awkward lengths (around tile and warp boundaries) via DOK row_ptr."""
    rng = np.random.default_rng(4)
    for rows in (1, 2, 31, 33, 2047, 2048, 2049, 4097, 100_003):
        n = rows * 2
        ri, ci = rng.integers(0, rows, n), rng.integers(0, 4, n)
        v = rng.integers(1, 5, n).astype(np.int32)
        got = S.CsrMatrix.from_triplets(rows, 4, ri, ci, v, handle=handle)
        off, idx, val = oracle.dok_to_csr(rows, 4, ri, ci, v)
        assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)


def test_cpp_host_mirror():
    """include/spam_csr.hpp (the C++ stand-in for the vendored spam_csr crate) over the C ABI."""
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "cpp", "test_mirror")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
