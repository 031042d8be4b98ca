"""Deterministic synthetic inputs for the configurations BASELINE.json names.

Host-side numpy only; these build the *inputs* of the hot path and are outside every timed region.
Every generator returns a csr.CsrMatrix-compatible tuple (rows, cols, offsets u64, indices u64, vals).
"""
from __future__ import annotations

import os

import numpy as np


def _csr_from_sorted_keys(rows: int, cols: int, keys: np.ndarray, vals: np.ndarray):
    """keys = row * cols + col, sorted ascending and unique."""
    r = (keys // np.uint64(cols)).astype(np.int64)
    idx = (keys % np.uint64(cols)).astype(np.uint64)
    counts = np.bincount(r, minlength=rows)
    offsets = np.zeros(rows + 1, dtype=np.uint64)
    np.cumsum(counts, out=offsets[1:])
    return rows, cols, offsets, idx, vals


def _sorted_unique(keys: np.ndarray) -> np.ndarray:
    """np.unique(keys) by sort + neighbour compare (numpy 2.3's hash-based unique takes 100 s on the 67 M keys
    of R-MAT 22; the AVX-512 sort takes one)."""
    if keys.shape[0] == 0:
        return keys
    s = np.sort(keys)
    keep = np.empty(s.shape[0], dtype=bool)
    keep[0] = True
    np.not_equal(s[1:], s[:-1], out=keep[1:])
    return s[keep]


def _nonzero_uniform(rng, n, dtype):
    v = rng.uniform(-1.0, 1.0, size=n)
    v[v == 0.0] = 0.5
    return v.astype(dtype)


def uniform_random(rows: int, cols: int, per_row: int, seed: int = 1, dtype=np.float64, int_range: int = 0):
    """C1 / C5 shape: `per_row` uniform column draws per row (with replacement, duplicates merged)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    r = np.repeat(np.arange(rows, dtype=np.uint64), per_row)
    c = rng.integers(0, cols, size=rows * per_row, dtype=np.uint64)
    keys = _sorted_unique(r * np.uint64(cols) + c)
    if int_range:
        v = rng.integers(-int_range, int_range + 1, size=keys.shape[0]).astype(dtype)
        v[v == 0] = 1
    else:
        v = _nonzero_uniform(rng, keys.shape[0], dtype)
    return _csr_from_sorted_keys(rows, cols, keys, v)


def poisson2d(n: int, dtype=np.float64, ny: int | None = None):
    """C2: 5-point Laplacian on an n x n grid (ny lines of n points when `ny` is given), row-major,
    diagonal 4, off-diagonals -1, sorted columns."""
    ny = n if ny is None else ny
    m = n * ny
    i = np.arange(m, dtype=np.int64)
    y, x = i // n, i % n
    cand = np.stack([i - n, i - 1, i, i + 1, i + n], axis=1)
    ok = np.stack([y > 0, x > 0, np.ones(m, bool), x < n - 1, y < ny - 1], axis=1)
    v = np.broadcast_to(np.array([-1, -1, 4, -1, -1], dtype=dtype), (m, 5))
    offsets = np.zeros(m + 1, dtype=np.uint64)
    np.cumsum(ok.sum(axis=1), out=offsets[1:])
    return m, m, offsets, cand[ok].astype(np.uint64), np.ascontiguousarray(v[ok])


def stencil27(n: int, dtype=np.float64):
    """C3: 27-point stencil on an n^3 grid, centre 26, others -1, sorted columns."""
    m = n * n * n
    d = np.array([-1, 0, 1], dtype=np.int64)
    dz, dy, dx = [a.ravel() for a in np.meshgrid(d, d, d, indexing="ij")]
    lin = dz * n * n + dy * n + dx                         # ascending in (dz, dy, dx) order
    w = np.where(lin == 0, 26, -1).astype(dtype)
    idx_parts, val_parts, cnt_parts = [], [], []
    for z0 in range(0, n, 8):                              # chunk by z-planes to bound host memory
        z1 = min(n, z0 + 8)
        i = np.arange(z0 * n * n, z1 * n * n, dtype=np.int64)
        z, y, x = i // (n * n), (i // n) % n, i % n
        ok = ((z[:, None] + dz >= 0) & (z[:, None] + dz < n) & (y[:, None] + dy >= 0) & (y[:, None] + dy < n) &
              (x[:, None] + dx >= 0) & (x[:, None] + dx < n))
        cand = i[:, None] + lin[None, :]
        idx_parts.append(cand[ok].astype(np.uint64))
        val_parts.append(np.broadcast_to(w, ok.shape)[ok])
        cnt_parts.append(ok.sum(axis=1))
    offsets = np.zeros(m + 1, dtype=np.uint64)
    np.cumsum(np.concatenate(cnt_parts), out=offsets[1:])
    return m, m, offsets, np.concatenate(idx_parts), np.ascontiguousarray(np.concatenate(val_parts))


def rmat(scale: int, edge_factor: int = 16, abcd=(0.45, 0.15, 0.15, 0.25), seed: int = 42, dtype=np.float64):
    """C4: R-MAT power-law graph, 2^scale rows, duplicates merged.  Default skew is the milder
    (0.45, 0.15, 0.15, 0.25): Graph500's (0.57, 0.19, 0.19, 0.05) makes A*A infeasible."""
    n = 1 << scale
    ne = edge_factor * n
    a, b, c, _ = abcd

    # Level l of edge e uses draw number l * ne + e of the PCG64 stream (one draw per double), so edge chunks
    # can be generated independently (cache-resident, one thread each) and still give the same matrix as one
    # pass per level over all the edges.
    def chunk(e0, e1):
        r = np.zeros(e1 - e0, dtype=np.uint64)
        col = np.zeros(e1 - e0, dtype=np.uint64)
        for level in range(scale):
            bg = np.random.PCG64(seed)
            bg.advance(level * ne + e0)
            u = np.random.Generator(bg).random(e1 - e0)
            rbit = (u >= a + b).astype(np.uint64)
            cbit = (((u >= a) & (u < a + b)) | (u >= a + b + c)).astype(np.uint64)
            r = (r << np.uint64(1)) | rbit
            col = (col << np.uint64(1)) | cbit
        return r * np.uint64(n) + col

    step = 1 << 18
    spans = [(e0, min(ne, e0 + step)) for e0 in range(0, ne, step)]
    if len(spans) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
            parts = list(ex.map(lambda sp: chunk(*sp), spans))
        allkeys = np.concatenate(parts)
    else:
        allkeys = chunk(0, ne)
    rng = np.random.Generator(np.random.PCG64(seed))
    rng.bit_generator.advance(scale * ne)
    keys = _sorted_unique(allkeys)
    v = _nonzero_uniform(rng, keys.shape[0], dtype)
    return _csr_from_sorted_keys(n, n, keys, v)


def transpose(mat):
    """Host transpose by key-swap sort (generator utility for C5's B = A^T)."""
    rows, cols, offsets, indices, vals = mat
    r = np.repeat(np.arange(rows, dtype=np.uint64), np.diff(offsets).astype(np.int64))
    keys = indices * np.uint64(rows) + r
    order = np.argsort(keys, kind="stable")
    return _csr_from_sorted_keys(cols, rows, keys[order], vals[order])


def triplets_with_rewrites(mat, seed: int = 5, dup_frac: float = 0.01, zero_frac: float = 0.001):
    """C5 DOK input: the matrix's entries as a shuffled triplet stream with ~dup_frac overwritten keys
    (an earlier write with another value) and ~zero_frac entries deleted again by a later zero write."""
    rows, cols, offsets, indices, vals = mat
    rng = np.random.Generator(np.random.PCG64(seed))
    nnz = indices.shape[0]
    r = np.repeat(np.arange(rows, dtype=np.uint64), np.diff(offsets).astype(np.int64))
    order = rng.permutation(nnz)
    tr, tc, tv = r[order], indices[order], vals[order].copy()
    ndup = int(nnz * dup_frac)
    nzero = int(nnz * zero_frac)
    # earlier writes that the final entry overwrites: prepend with different values
    pick = rng.choice(nnz, size=ndup, replace=False)
    er, ec = tr[pick], tc[pick]
    ev = (tv[pick] * 3 + 1).astype(vals.dtype)
    ev[ev == 0] = 7
    # later zero writes that delete entries: append
    zpick = rng.choice(nnz, size=nzero, replace=False)
    zr, zc = tr[zpick], tc[zpick]
    zv = np.zeros(nzero, dtype=vals.dtype)
    return (np.concatenate([er, tr, zr]), np.concatenate([ec, tc, zc]), np.concatenate([ev, tv, zv]))


def spgemm_counts(a, b):
    """(flops P, per-row flops) of A*B: P = sum over A entries of nnz(B row)  (mul_hash.rs:39-50)."""
    _, _, ao, ai, _ = a
    _, _, bo, _, _ = b
    blen = np.diff(bo).astype(np.int64)
    per_entry = blen[ai.astype(np.int64)]
    cs = np.concatenate([[0], np.cumsum(per_entry)])
    per_row = cs[ao.astype(np.int64)[1:]] - cs[ao.astype(np.int64)[:-1]]
    return int(per_entry.sum()), per_row


def algorithmic_bytes_spgemm(rows, nnz_a, flops, nnz_c, val_size):
    """read A once, one B entry per product, write C once (u32 idx, u64 ptr)."""
    return nnz_a * (4 + val_size) + (rows + 1) * 8 + flops * (4 + val_size) + nnz_c * (4 + val_size) + (rows + 1) * 8


def algorithmic_bytes_spmv(rows, cols, nnz, val_size):
    return nnz * (4 + val_size) + (rows + 1) * 8 + cols * val_size + rows * val_size


WORKLOADS = {
    "uniform10k": lambda: uniform_random(10_000, 10_000, 10, seed=1),
    "poisson2048": lambda: poisson2d(2048),
    "stencil160": lambda: stencil27(160),
    "rmat22": lambda: rmat(22),
}
