#!/usr/bin/env python
"""Launch-list probe: Poisson 2048^2 A*A with every row's entries shuffled (CsrMatrix<T,false> inputs), three products.
Run under `ncu --metrics gpu__time_duration.sum` to see which kernels the unsorted-input line of bench.py pays for."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_matrix_b200 as S  # noqa: E402
from sparse_matrix_b200 import generators as G  # noqa: E402

mat = G.poisson2d(2048)
rows, cols = mat[0], mat[1]
rng = np.random.default_rng(7)
lens = np.diff(mat[2].astype(np.int64))
keyr = np.repeat(np.arange(rows), lens) + rng.random(len(mat[3]))
perm = np.argsort(keyr, kind="stable")
h = S.Handle(0)
dU = S.DeviceCsr.upload(S.CsrMatrix(rows, cols, mat[4][perm], mat[3][perm], mat[2], is_sorted=False), h)
for _ in range(3):
    dU.matmul(dU).free()
h.synchronize()
print("stats", h.stats()["num_bin_rows"])
