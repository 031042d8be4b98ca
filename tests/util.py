"""Shared helpers for the parity tests."""
import numpy as np

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}  # BASELINE.json north_star tolerances


def random_csr(rng, rows, cols, row_nnz, dtype=np.float64, sorted_rows=True, int_range=50, zero_frac=0.0):
    """row_nnz: int (uniform draws per row, duplicates merged) or array of per-row targets."""
    if np.isscalar(row_nnz):
        row_nnz = np.full(rows, int(row_nnz))
    row_nnz = np.minimum(np.asarray(row_nnz, dtype=np.int64), cols)
    idx_parts, offsets = [], np.zeros(rows + 1, dtype=np.uint64)
    for r in range(rows):
        k = int(row_nnz[r])
        if k > cols // 2:
            c = rng.permutation(cols)[:k]
        else:
            c = np.unique(rng.integers(0, cols, size=k))
        c = np.sort(c) if sorted_rows else rng.permutation(c)
        idx_parts.append(c.astype(np.uint64))
        offsets[r + 1] = offsets[r] + np.uint64(len(c))
    indices = np.concatenate(idx_parts) if idx_parts else np.empty(0, np.uint64)
    n = indices.shape[0]
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        vals = rng.uniform(-1, 1, size=n).astype(dtype)
        vals[vals == 0] = 0.25
    else:
        vals = rng.integers(-int_range, int_range + 1, size=n).astype(dtype)
        vals[vals == 0] = 1
    if zero_frac:
        vals[rng.random(n) < zero_frac] = 0  # explicit zeros propagate
    return rows, cols, offsets, indices, vals


def as_csr_matrix(t, is_sorted=True):
    from sparse_matrix_b200 import CsrMatrix
    r, c, o, i, v = t
    return CsrMatrix(r, c, v, i, o, is_sorted=is_sorted)


def check_against_oracle(oracle, a, b, c, exact_values=False):
    """c: CsrMatrix from the GPU.  Structure bit-exact against mul_hash::<_, true>; integer values
    bit-exact; floats within TOL * sum|products| per entry (the oracle run on |A|, |B|)."""
    off, idx, val = oracle.mul_hash(a, b, True)
    assert c.invariants()
    assert np.array_equal(c.offsets, off), "row_ptr differs"
    assert np.array_equal(c.indices, idx), "col_idx differs"
    dt = np.dtype(val.dtype)
    if dt.kind != "f" or exact_values:
        assert np.array_equal(c.vals.view(np.uint8), val.view(np.uint8)) or np.array_equal(c.vals, val), "values differ"
    else:
        aa = (a[0], a[1], a[2], a[3], np.abs(a[4]))
        bb = (b[0], b[1], b[2], b[3], np.abs(b[4]))
        _, _, sabs = oracle.mul_hash(aa, bb, True)
        err = np.abs(c.vals.astype(np.float64) - val.astype(np.float64))
        bound = TOL[dt] * sabs.astype(np.float64)
        bad = err > bound
        assert not bad.any(), f"{bad.sum()} values out of tolerance; worst {np.max(err / np.maximum(sabs, 1e-300))}"
    return off, idx, val
