#!/usr/bin/env python
"""Development probe for ncu: the product of ONE shard (r of n, cost-balanced) of a workload, 3 times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_b200 as S
from bench import make_workload
wl, r, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
h = S.Handle(0)
mat = make_workload(wl)
dA = S.DeviceCsr.upload(S.CsrMatrix(mat[0], mat[1], mat[4], mat[3], mat[2]), h)
starts, _ = dA.rows_to_parts(dA, n, balance="cost")
blk = dA.slice_rows(int(starts[r]), int(starts[r + 1]))
h.set_timing(True)
for _ in range(3):
    c = blk.matmul(dA); s = h.stats(); c.free()
print(s)
