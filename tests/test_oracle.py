"""CPU tests pinning the oracle (oracle/spam_oracle.cpp) before it is trusted as the checker.

The reference holds no golden vectors for this path; what it holds are properties,
re-created here against the restatement:
  * spam_csr/src/tests.rs:356-371   mul_hash == dense DokMatrix product, Wrapping<i8>, dims <= 4,
                                    inputs with shuffled (unsorted) rows
  * spam_csr/src/mul_hash.rs:204-224 rows_to_threads: flop.len()==rows, offsets sorted, last==rows
  * fuzz/fuzz_targets/mul_hash.rs    invariants always; Higham (3.13) bound when l*m*n < 2^15
plus hand-derived known-answer vectors (tests/golden/linprobe_kat.json) and scipy as an independent
structural cross-check.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
from hypothesis import given, settings
from hypothesis import strategies as st

from util import random_csr

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "linprobe_kat.json")))


# ---------------------------------------------------------------------------------------------
# known answers
# ---------------------------------------------------------------------------------------------
def test_table_size_rule(oracle):
    for cap, want in GOLD["table_size_for"].items():
        assert oracle.table_size_for(int(cap)) == want


def test_hash_slots(oracle):
    for key, slot in GOLD["slot_cap16"].items():
        assert oracle.hash_of(int(key)) & 15 == slot
    for key, slot in GOLD["slot_cap32"].items():
        assert oracle.hash_of(int(key)) & 31 == slot
    assert oracle.hash_of(0xFFFFFFFE) == (0xFFFFFFFE * 107) % (1 << 32)  # wrapping_mul


@pytest.mark.parametrize("name", ["set_collision", "set_growth"])
def test_hashset_kat(oracle, name):
    g = GOLD[name]
    n, ub, slots = oracle.hashset_run(g["keys"])
    assert (n, ub) == (g["len"], g["upper_bound"])
    occ = {str(i): int(k) for i, k in enumerate(slots) if k != 0xFFFFFFFF}
    assert occ == g["occupied"]


def test_map_slot_order_kat(oracle):
    g = GOLD["map_slot_order"]
    a = (g["a"]["rows"], g["a"]["cols"], g["a"]["offsets"], g["a"]["indices"], np.array(g["a"]["vals"]))
    b = (g["b"]["rows"], g["b"]["cols"], g["b"]["offsets"], g["b"]["indices"], np.array(g["b"]["vals"]))
    for mode, key in ((False, "unsorted"), (True, "sorted")):
        off, idx, val = oracle.mul_hash(a, b, mode)
        assert off.tolist() == g[key]["offsets"]
        assert idx.tolist() == g[key]["indices"]
        assert val.tolist() == g[key]["vals"]


@pytest.mark.parametrize("name", ["map_reuse_shrink", "map_collision_chain"])
def test_map_reuse_and_collision_chain_kats(oracle, name):
    """One map serves every row of a thread block and is re-sized per row (map.rs:49-58); entry() probes with
    wrap-around and accumulates on a hit (map.rs:66-93): hand-derived slot orders and sums."""
    g = GOLD[name]
    a = (g["a"]["rows"], g["a"]["cols"], g["a"]["offsets"], g["a"]["indices"], np.array(g["a"]["vals"]))
    if g["b"] == "identity 65":
        b = (65, 65, np.arange(66), np.arange(65), np.ones(65))
    else:
        b = (g["b"]["rows"], g["b"]["cols"], g["b"]["offsets"], g["b"]["indices"], np.array(g["b"]["vals"]))
    for tnum in (1, 2, 0):            # the answer does not depend on how rows are split over threads
        for mode, key in ((False, "unsorted"), (True, "sorted")):
            off, idx, val = oracle.mul_hash(a, b, mode, tnum)
            assert off.tolist() == g[key]["offsets"]
            assert idx.tolist() == g[key]["indices"], (name, key, tnum)
            assert val.tolist() == g[key]["vals"]


def test_set_grow_after_shrink_kat(oracle):
    g = GOLD["set_grow_after_shrink"]
    n, ub, alloc, slots = oracle.hashset_run2(g["keys"], g["initial_capacity"], g["shrink_to"])
    assert (n, ub, alloc) == (g["len"], g["upper_bound"], 128)          # no re-allocation: 128 slots from with_capacity(40)
    occ = {str(i): int(k) for i, k in enumerate(slots) if k != 0xFFFFFFFF}
    assert occ == g["occupied"]
    n, ub, slots = oracle.hashset_run(g["keys"])                         # HashSet::new(): grow() re-allocates 16 -> 32
    assert (n, ub) == (g["len"], g["upper_bound"])
    assert {str(i): int(k) for i, k in enumerate(slots) if k != 0xFFFFFFFF} == g["occupied"]


def test_rows_to_threads_ties_kat(oracle):
    g = GOLD["rows_to_threads_ties"]
    for tnum, want in g["rows_offset"].items():
        flop, ro = oracle.rows_to_threads(g["a"]["rows"], g["a"]["offsets"], g["a"]["indices"], g["b_offsets"], int(tnum))
        assert flop.tolist() == g["flops"] and ro.tolist() == want, (tnum, ro.tolist())


def test_unfused_cancellation_kat(oracle):
    g = GOLD["unfused_cancellation"]
    av = np.array([float.fromhex(x) for x in g["a"]["vals_hex"]])
    bv = np.array([float.fromhex(x) for x in g["b"]["vals_hex"]])
    a = (1, 2, g["a"]["offsets"], g["a"]["indices"], av)
    b = (2, 8, g["b"]["offsets"], g["b"]["indices"], bv)
    off, idx, val = oracle.mul_hash(a, b, True)
    assert off.tolist() == g["c"]["offsets"] and idx.tolist() == g["c"]["indices"]
    assert val.tolist() == [0.0]  # exactly zero (no FMA) and kept as an explicit entry


# ---------------------------------------------------------------------------------------------
# reference property: mul_hash == dense DOK product (tests.rs:356-371)
# ---------------------------------------------------------------------------------------------
def _dok_to_unsorted_csr(dense, rng):
    """CsrMatrix::from_dok(shuffle) (lib.rs:337-358): entries of the DOK (non-zeros), rows shuffled."""
    rows, cols = dense.shape
    offsets, idx, val = [0], [], []
    for r in range(rows):
        c = np.nonzero(dense[r])[0]
        c = rng.permutation(c)
        idx.extend(c.tolist())
        val.extend(dense[r, c].tolist())
        offsets.append(len(idx))
    return rows, cols, np.array(offsets, np.uint64), np.array(idx, np.uint64), np.array(val, dense.dtype)


def _csr_to_dense_dropping_zeros(rows, cols, off, idx, val):
    d = np.zeros((rows, cols), dtype=val.dtype)
    for r in range(rows):
        for e in range(int(off[r]), int(off[r + 1])):
            d[r, int(idx[e])] = val[e]  # set_element: zero => absent, which a dense array shows as 0 too
    return d


small = st.integers(min_value=1, max_value=4)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(l=small, m=small, n=small, data=st.data())
def test_mul_hash_commutes_with_dense_dok_i8(oracle, l, m, n, data):
    # arb_fixed_size_matrix: up to 2*rows*cols random set_element calls (spam_dok lib.rs:244-260)
    def arb(rows, cols):
        d = np.zeros((rows, cols), dtype=np.int8)
        k = data.draw(st.integers(0, 2 * rows * cols))
        for _ in range(k):
            d[data.draw(st.integers(0, rows - 1)), data.draw(st.integers(0, cols - 1))] = data.draw(
                st.integers(-128, 127))
        return d
    da, db = arb(l, m), arb(m, n)
    rng = np.random.default_rng(data.draw(st.integers(0, 2**31)))
    a, b = _dok_to_unsorted_csr(da, rng), _dok_to_unsorted_csr(db, rng)
    want = oracle.dok_dense_mul(da, db)          # wrapping i8 triple loop
    assert np.array_equal(want, (da.astype(np.int64) @ db.astype(np.int64)).astype(np.int8))  # restated loop sanity
    for mode in (False, True):
        off, idx, val = oracle.mul_hash(a, b, mode)
        _assert_invariants(l, n, off, idx, val, mode)
        assert np.array_equal(_csr_to_dense_dropping_zeros(l, n, off, idx, val), want)


def _assert_invariants(rows, cols, off, idx, val, is_sorted):
    assert len(idx) == len(val)                       # invariant1
    assert len(off) == rows + 1                       # invariant2
    assert np.all(off[1:] >= off[:-1])                # invariant3
    assert int(off[rows]) == len(idx)                 # invariant4
    assert (len(idx) == 0) or int(idx.max()) < cols   # invariant5
    assert int(off[0]) == 0                           # invariant7
    for r in range(rows):                             # invariant6
        seg = idx[int(off[r]):int(off[r + 1])]
        if is_sorted:
            assert np.all(seg[1:] > seg[:-1])
        else:
            assert len(np.unique(seg)) == len(seg)


# ---------------------------------------------------------------------------------------------
# reference property: rows_to_threads (mul_hash.rs:204-224)
# ---------------------------------------------------------------------------------------------
@settings(max_examples=100, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**31), rows=st.integers(1, 40), inner=st.integers(1, 40), cols=st.integers(1, 40),
       tnum=st.integers(1, 16))
def test_rows_to_threads_shape(oracle, seed, rows, inner, cols, tnum):
    rng = np.random.default_rng(seed)
    a = random_csr(rng, rows, inner, rng.integers(0, inner + 1, size=rows))
    b = random_csr(rng, inner, cols, rng.integers(0, cols + 1, size=inner))
    flop, ro = oracle.rows_to_threads(rows, a[2], a[3], b[2], tnum)
    assert len(flop) == rows
    assert np.all(ro[1:] >= ro[:-1])
    assert int(ro[-1]) == rows and int(ro[0]) == 0 and len(ro) == tnum + 1
    blen = np.diff(b[2]).astype(np.int64)
    want = np.array([blen[a[3][int(a[2][r]):int(a[2][r + 1])].astype(np.int64)].sum() for r in range(rows)])
    assert np.array_equal(flop.astype(np.int64), want)


# ---------------------------------------------------------------------------------------------
# fuzz target property: invariants + Higham bound (fuzz_targets/mul_hash.rs, spam_dok lib.rs:56-92)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(25))
def test_higham_bound_f64(oracle, seed):
    rng = np.random.default_rng(seed)
    l, m, n = (int(x) for x in rng.integers(1, 32, size=3))
    a = random_csr(rng, l, m, rng.integers(0, m + 1, size=l), sorted_rows=False)
    b = random_csr(rng, m, n, rng.integers(0, n + 1, size=m), sorted_rows=False)
    # wide dynamic range, like arbitrary f64
    a = a[:4] + (a[4] * np.exp(rng.uniform(-20, 20, size=a[4].shape)),)
    b = b[:4] + (b[4] * np.exp(rng.uniform(-20, 20, size=b[4].shape)),)
    off, idx, val = oracle.mul_hash(a, b, False)
    _assert_invariants(l, n, off, idx, val, False)
    A = sp.csr_matrix((a[4], a[3].astype(np.int64), a[2].astype(np.int64)), shape=(l, m)).toarray()
    B = sp.csr_matrix((b[4], b[3].astype(np.int64), b[2].astype(np.int64)), shape=(m, n)).toarray()
    got = _csr_to_dense_dropping_zeros(l, n, off, idx, val)
    expected = oracle.dok_dense_mul(A, B)
    u = np.finfo(np.float64).eps / 2
    nn = float(n)                                   # `self.cols()` in is_good_approx_of_mul
    gamma = nn * u / (1 - nn * u)
    inf = lambda M: np.abs(M).sum(axis=1).max() if M.size else 0.0
    assert inf(expected - got) <= 2 * gamma * inf(A) * inf(B)


# ---------------------------------------------------------------------------------------------
# independent cross-checks
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64])
@pytest.mark.parametrize("seed", [0, 1])
def test_structure_matches_scipy(oracle, dtype, seed):
    rng = np.random.default_rng(seed)
    a = random_csr(rng, 300, 200, rng.integers(0, 30, size=300), dtype=dtype)
    b = random_csr(rng, 200, 400, rng.integers(0, 50, size=200), dtype=dtype)
    off, idx, val = oracle.mul_hash(a, b, True)
    A = sp.csr_matrix((np.ones(len(a[4])), a[3].astype(np.int64), a[2].astype(np.int64)), shape=(300, 200))
    B = sp.csr_matrix((np.ones(len(b[4])), b[3].astype(np.int64), b[2].astype(np.int64)), shape=(200, 400))
    Cp = (A @ B).tocsr()
    Cp.sort_indices()
    assert np.array_equal(off.astype(np.int64), Cp.indptr)       # pattern(A)*pattern(B): zeros kept
    assert np.array_equal(idx.astype(np.int64), Cp.indices)
    if np.dtype(dtype).kind == "f":
        Av = sp.csr_matrix((a[4].astype(np.float64), a[3].astype(np.int64), a[2].astype(np.int64)), shape=(300, 200))
        Bv = sp.csr_matrix((b[4].astype(np.float64), b[3].astype(np.int64), b[2].astype(np.int64)), shape=(200, 400))
        dense = (Av @ Bv).toarray()
        rows = np.repeat(np.arange(300), np.diff(off).astype(np.int64))
        tol = 1e-12 if dtype == np.float64 else 1e-4
        assert np.allclose(val, dense[rows, idx.astype(np.int64)], rtol=tol, atol=tol)


def test_thread_count_does_not_change_bits(oracle):
    """Table capacity depends only on nnz_i (mul_hash.rs:144), so the output is identical for any tnum."""
    rng = np.random.default_rng(7)
    a = random_csr(rng, 500, 500, rng.integers(0, 40, size=500), sorted_rows=False)
    ref = [oracle.mul_hash(a, a, mode, tnum=1) for mode in (False, True)]
    for tnum in (2, 3, 8, 64):
        for mode in (False, True):
            got = oracle.mul_hash(a, a, mode, tnum=tnum)
            for x, y in zip(ref[int(mode)], got):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_sorted_is_rowwise_sort_of_unsorted(oracle):
    rng = np.random.default_rng(3)
    a = random_csr(rng, 200, 300, 12, sorted_rows=False)
    b = random_csr(rng, 300, 250, 9, sorted_rows=False)
    ou, iu, vu = oracle.mul_hash(a, b, False)
    os_, is_, vs = oracle.mul_hash(a, b, True)
    assert np.array_equal(ou, os_)
    for r in range(200):
        lo, hi = int(ou[r]), int(ou[r + 1])
        order = np.argsort(iu[lo:hi])
        assert np.array_equal(iu[lo:hi][order], is_[lo:hi])
        assert np.array_equal(vu[lo:hi][order], vs[lo:hi])


def test_explicit_input_zeros_propagate(oracle):
    a = (1, 2, [0, 2], [0, 1], np.array([0.0, 2.0]))
    b = (2, 3, [0, 1, 2], [2, 0], np.array([5.0, 0.0]))
    off, idx, val = oracle.mul_hash(a, b, True)
    assert off.tolist() == [0, 2] and idx.tolist() == [0, 2] and val.tolist() == [0.0, 0.0]


def test_integer_wrapping(oracle):
    big = np.int64(2**62)
    a = (1, 2, [0, 2], [0, 1], np.array([big, big], dtype=np.int64))
    b = (2, 1, [0, 1, 2], [0, 0], np.array([2, 2], dtype=np.int64))
    off, idx, val = oracle.mul_hash(a, b, True)
    assert val.tolist() == [0]  # 2^63 + 2^63 wraps to 0; the entry stays
    a8 = (1, 1, [0, 1], [0], np.array([100], dtype=np.int8))
    b8 = (1, 1, [0, 1], [0], np.array([3], dtype=np.int8))
    assert oracle.mul_hash(a8, b8, True)[2].tolist() == [np.int8(300 - 256)]


# ---------------------------------------------------------------------------------------------
# DOK -> CSR (spam_dok lib.rs:167-176 + spam_csr lib.rs:315-334) and SpMV restatement
# ---------------------------------------------------------------------------------------------
def test_dok_last_write_wins_and_zero_deletes(oracle):
    ri = [2, 0, 2, 0, 3, 3, 0]
    ci = [1, 4, 1, 4, 0, 0, 2]
    v = np.array([5.0, 1.0, 6.0, 0.0, 7.0, 8.0, 9.0])   # (2,1): 5 then 6; (0,4): 1 then deleted; (3,0): 7 then 8
    off, idx, val = oracle.dok_to_csr(5, 5, ri, ci, v)
    assert off.tolist() == [0, 1, 1, 2, 3, 3]            # rows 1 and 4 empty: repeated offsets
    assert idx.tolist() == [2, 1, 0] and val.tolist() == [9.0, 6.0, 8.0]
    # delete then re-insert keeps the later value
    off, idx, val = oracle.dok_to_csr(1, 1, [0, 0, 0], [0, 0, 0], np.array([1.0, 0.0, 2.0]))
    assert val.tolist() == [2.0]
    with pytest.raises(IndexError):
        oracle.dok_to_csr(2, 2, [2], [0], np.array([1.0]))


def test_dok_matches_python_dict_model(oracle):
    rng = np.random.default_rng(11)
    n = 5000
    ri, ci = rng.integers(0, 60, n), rng.integers(0, 70, n)
    v = rng.integers(-2, 3, n).astype(np.int64)
    model = {}
    for r, c, t in zip(ri, ci, v):
        if t == 0:
            model.pop((int(r), int(c)), None)
        else:
            model[(int(r), int(c))] = int(t)
    off, idx, val = oracle.dok_to_csr(60, 70, ri, ci, v)
    keys = sorted(model)
    assert idx.tolist() == [k[1] for k in keys] and val.tolist() == [model[k] for k in keys]
    counts = np.bincount([k[0] for k in keys], minlength=60)
    assert np.array_equal(np.diff(off).astype(np.int64), counts)


def test_spmv_is_mul_hash_with_column_vector(oracle):
    rng = np.random.default_rng(5)
    a = random_csr(rng, 80, 60, rng.integers(0, 20, size=80), sorted_rows=False)
    x = rng.uniform(-1, 1, size=60)
    y = oracle.spmv(80, 60, a[2], a[3], a[4], x)
    xm = (60, 1, np.arange(61, dtype=np.uint64), np.zeros(60, np.uint64), x)   # one explicit entry per k
    off, idx, val = oracle.mul_hash(a, xm, True)
    dense = np.zeros(80)
    rows = np.repeat(np.arange(80), np.diff(off).astype(np.int64))
    dense[rows] = val
    assert np.array_equal(dense, y)   # bit-identical: same order, same unfused arithmetic


@pytest.mark.parametrize("dtype", [np.float64, np.int32])
@pytest.mark.parametrize("sorted_rows", [True, False])
def test_transpose_literal_loops_equal_stable_column_sort(oracle, dtype, sorted_rows):
    """CsrMatrix::transpose (lib.rs:256-264): the reference's own (j, i) double loop of set_element calls
    against the stable counting sort the GPU path is checked with; explicit zeros are kept; an independent
    check against scipy; (A^T)^T is A with sorted rows."""
    rng = np.random.default_rng(11)
    for rows, cols, deg in ((1, 1, 1), (7, 3, 2), (5, 40, 9), (40, 5, 4), (33, 33, 0), (24, 31, 12)):
        a = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, sorted_rows=sorted_rows,
                       zero_frac=0.2)
        lit = oracle.transpose(a, literal=True)
        fast = oracle.transpose(a)
        for x, y in zip(lit, fast):
            assert np.array_equal(x, y)
        t_off, t_idx, t_val = fast
        assert int(t_off[-1]) == len(a[3]) and len(t_off) == cols + 1                  # zeros kept
        for j in range(cols):
            seg = t_idx[int(t_off[j]):int(t_off[j + 1])]
            assert np.all(np.diff(seg.astype(np.int64)) > 0)                          # rows sorted
        # structure through scipy on a zero-free copy of the values (scipy would be free to drop zeros)
        ones = np.arange(1, len(a[3]) + 1, dtype=np.float64)
        m = sp.csr_matrix((ones, a[3].astype(np.int64), a[2].astype(np.int64)), shape=(rows, cols)).T.tocsr()
        m.sort_indices()
        assert np.array_equal(m.indptr, t_off.astype(np.int64)) and np.array_equal(m.indices, t_idx.astype(np.int64))
        assert np.array_equal(a[4][(m.data - 1).astype(np.int64)], t_val)
        back = oracle.transpose((cols, rows) + fast)
        srt = oracle.transpose((cols, rows) + oracle.transpose(a))                     # = A with sorted rows
        assert np.array_equal(back[0], a[2]) and np.array_equal(back[0], srt[0])
        if sorted_rows:
            assert np.array_equal(back[1], a[3]) and np.array_equal(back[2], a[4])


@pytest.mark.parametrize("op", ["add", "sub"])
def test_elementwise_matches_dense_and_keeps_zeros(oracle, op):
    """apply_elementwise (lib.rs:83-149): against dense numpy on the union pattern; cancellation zeros and
    explicit zeros stay; the IS_SORTED = false branch gives the same entries except that a -0.0 only in the
    left operand keeps its sign under add; i8 wraps (the reference's test type, tests.rs:334-354)."""
    rng = np.random.default_rng(17)
    for dtype in (np.float64, np.int32, np.int8):
        for rows, cols, deg in ((1, 1, 1), (6, 9, 4), (30, 17, 8), (12, 40, 0)):
            a = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, zero_frac=0.15)
            b = random_csr(rng, rows, cols, rng.integers(0, deg + 1, size=rows), dtype=dtype, zero_frac=0.15)
            off, idx, val = oracle.ewise(a, b, op, True)
            da, db = np.zeros((rows, cols), dtype), np.zeros((rows, cols), dtype)
            pat = np.zeros((rows, cols), bool)
            for m, d in ((a, da), (b, db)):
                for r in range(rows):
                    for e in range(int(m[2][r]), int(m[2][r + 1])):
                        d[r, int(m[3][e])] = m[4][e]
                        pat[r, int(m[3][e])] = True
            with np.errstate(over="ignore"):
                want = (da + db) if op == "add" else (da - db)
            assert int(off[-1]) == int(pat.sum())                      # the union pattern, nothing dropped
            for r in range(rows):
                seg = idx[int(off[r]):int(off[r + 1])].astype(np.int64)
                assert np.array_equal(seg, np.nonzero(pat[r])[0])      # sorted by column
                assert np.array_equal(val[int(off[r]):int(off[r + 1])], want[r, seg])
            # unsorted branch on row-shuffled operands: same entries
            off2, idx2, val2 = oracle.ewise(a, b, op, False)
            assert np.array_equal(off, off2) and np.array_equal(idx, idx2) and np.array_equal(val, val2)
    # the one observable difference between the branches: left-only -0.0 under add
    a = (1, 2, [0, 1], [0], np.array([-0.0]))
    b = (1, 2, [0, 1], [1], np.array([2.0]))
    assert np.signbit(oracle.ewise(a, b, "add", False)[2][0]) and not np.signbit(oracle.ewise(a, b, "add", True)[2][0])
    assert np.signbit(oracle.ewise(a, b, "sub", True)[2][0])
    with pytest.raises(ValueError):
        oracle.ewise(a, (2, 2, [0, 0, 0], [], np.array([])), "add")


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), rows=st.integers(1, 9), inner=st.integers(1, 9), cols=st.integers(1, 9))
def test_algebraic_identities_between_the_restatements(oracle, seed, rows, inner, cols):
    """Cross-checks between the independent restatements (integers, so everything is exact):
    (A B)^T = B^T A^T,  (A + B)^T = A^T + B^T,  (A + B) - B has A's values on the union pattern,
    A (B + C) = A B + A C as dense matrices (patterns may differ by explicit zeros)."""
    rng = np.random.default_rng(seed)

    def dense(m):
        r, c, off, idx, val = m
        d = np.zeros((r, c), np.int64)
        for i in range(r):
            for e in range(int(off[i]), int(off[i + 1])):
                d[i, int(idx[e])] = val[e]
        return d
    a = random_csr(rng, rows, inner, rng.integers(0, inner + 1, size=rows), dtype=np.int64, int_range=9)
    b = random_csr(rng, inner, cols, rng.integers(0, cols + 1, size=inner), dtype=np.int64, int_range=9)
    c = random_csr(rng, inner, cols, rng.integers(0, cols + 1, size=inner), dtype=np.int64, int_range=9)
    T = lambda m: (m[1], m[0]) + oracle.transpose(m)
    ab = (rows, cols) + oracle.mul_hash(a, b, True)
    btat = (cols, rows) + oracle.mul_hash(T(b), T(a), True)
    abt = T(ab)
    assert np.array_equal(abt[2], btat[2]) and np.array_equal(abt[3], btat[3]) and np.array_equal(abt[4], btat[4])
    bc = (inner, cols) + oracle.ewise(b, c, "add", True)
    lhs = T(bc)
    rhs = oracle.ewise(T(b), T(c), "add", True)
    assert all(np.array_equal(x, y) for x, y in zip(lhs[2:], rhs))
    back = oracle.ewise(bc, c, "sub", True)
    assert np.array_equal(back[0], bc[2]) and np.array_equal(back[1], bc[3])            # union pattern kept
    assert np.array_equal(dense((inner, cols) + back), dense(b))
    left = dense((rows, cols) + oracle.mul_hash(a, bc, True))
    right = dense((rows, cols) + oracle.ewise(ab, (rows, cols) + oracle.mul_hash(a, c, True), "add", True))
    assert np.array_equal(left, right)
