#!/usr/bin/env python
"""Development probe: one matrix, the product timed under each SPAM_LANES mode (0 = single stream)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_matrix_b200 as S
from bench import make_workload
wl = sys.argv[1]
h = S.Handle(0)
mat = make_workload(wl)
dA = S.DeviceCsr.upload(S.CsrMatrix(mat[0], mat[1], mat[4], mat[3], mat[2]), h)
h.set_timing(True)
for mode in sys.argv[2:]:
    os.environ["SPAM_LANES"] = mode
    best = None
    for _ in range(5):
        c = dA.matmul(dA); s = h.stats(); c.free()
        if best is None or s["ms_total"] < best["ms_total"]: best = s
    print(json.dumps({"workload": wl, "lanes": mode, "ms_total": round(best["ms_total"], 3), "sym": round(best["ms_symbolic"], 3),
                      "num": round(best["ms_numeric"], 3)}), flush=True)
starts, _ = dA.rows_to_parts(dA, 8, balance="cost")
for r in (0, 3, 7):
    blk = dA.slice_rows(int(starts[r]), int(starts[r + 1]))
    for mode in sys.argv[2:]:
        os.environ["SPAM_LANES"] = mode
        best = None
        for _ in range(5):
            c = blk.matmul(dA); s = h.stats(); c.free()
            if best is None or s["ms_total"] < best["ms_total"]: best = s
        print(json.dumps({"shard": r, "of": 8, "lanes": mode, "ms_total": round(best["ms_total"], 3), "sym": round(best["ms_symbolic"], 3),
                          "num": round(best["ms_numeric"], 3)}), flush=True)
    blk.free()
