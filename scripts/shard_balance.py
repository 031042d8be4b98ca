#!/usr/bin/env python
"""Load balance of the row partition, measured on ONE GPU: the product has no data-path collective, so
the time rank r of N would take is the time of shard r run alone.  Prints, per shard, rows / flops /
nnz(C) / ms and the phase split, then max / mean (the strong-scaling efficiency the partition allows).

usage: shard_balance.py WORKLOAD N [N ...]"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import sparse_matrix_b200 as S  # noqa: E402
from bench import make_workload  # noqa: E402


BALANCE = os.environ.get("BALANCE", "cost")


def main():
    wl = sys.argv[1]
    parts = [int(x) for x in sys.argv[2:]] or [8]
    h = S.Handle(0)
    mat = make_workload(wl)
    A = S.CsrMatrix(mat[0], mat[1], mat[4], mat[3], mat[2])
    dA = S.DeviceCsr.upload(A, h)
    h.set_timing(True)

    def run(blk, reps=3):
        best, st = None, None
        for _ in range(reps + 1):
            c = blk.matmul(dA)
            s = h.stats()
            c.free()
            if best is None or s["ms_total"] < best:
                best, st = s["ms_total"], s
        return best, st

    whole, st = run(dA)
    print(json.dumps({"workload": wl, "whole_ms": whole, "flops": st["flops"], "nnz_c": st["nnz_c"]}), flush=True)
    for n in parts:
        starts, total = dA.rows_to_parts(dA, n, balance=BALANCE)
        ms = []
        for r in range(n):
            blk = dA.slice_rows(int(starts[r]), int(starts[r + 1]))
            t, s = run(blk)
            ms.append(t)
            print(json.dumps({"n": n, "shard": r, "rows": int(starts[r + 1] - starts[r]), "flops": s["flops"],
                              "nnz_c": s["nnz_c"], "ms": round(t, 3), "sym": round(s["ms_flop"] + s["ms_symbolic"], 3),
                              "num": round(s["ms_numeric"], 3), "num_bins": s["num_bin_rows"][:12]}), flush=True)
            blk.free()
        print(json.dumps({"n": n, "max_ms": max(ms), "mean_ms": float(np.mean(ms)), "speedup_vs_whole": whole / max(ms),
                          "balance": float(np.mean(ms)) / max(ms)}), flush=True)
    dA.free()
    h.close()


if __name__ == "__main__":
    main()
