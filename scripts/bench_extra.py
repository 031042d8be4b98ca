#!/usr/bin/env python
"""Extra measurements for the SURVEY §8 rows that bench.py's headline line does not cover:
SpMV (a7), DOK->CSR (a8) and the C5 rectangular i64 product, each with its HBM roofline fraction
(algorithmic bytes of SURVEY §8d / CUDA-event time / measured copy peak) and the CPU oracle beside it.
One JSON line per measurement.  Device-resident, CUDA events on the library's stream."""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import sparse_matrix_b200 as S  # noqa: E402
from bench import measured_peak  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from sparse_matrix_b200 import generators as G  # noqa: E402


_H = None


def timed(fn, stream, reps=20, warm=3):
    """ms per call: `reps` back-to-back calls on the handle's own stream between two stream syncs (host
    clock; the calls are asynchronous apart from the library's own syncs, so this is device time plus the
    launch gaps a caller would see)."""
    for _ in range(warm):
        fn()
    _H.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    _H.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


def main():
    which = sys.argv[1:] or ["spmv", "ewise", "dok", "rect"]
    dev = torch.device("cuda", 0)
    stream = None
    h = S.Handle(0)   # the handle's own non-blocking stream
    global _H
    _H = h
    L = h.L
    peak, src = measured_peak()

    if "spmv" in which:
        p = G.poisson2d(2048)
        A = S.CsrMatrix(p[0], p[1], p[4], p[3], p[2])
        dA = S.DeviceCsr.upload(A, h)
        rng = np.random.Generator(np.random.PCG64(2))
        x = rng.uniform(-1, 1, size=p[1])
        dx = torch.from_numpy(x).to(dev)
        dy = torch.empty(p[0], dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        ms = timed(lambda: S._lib.check(h.h, L.spam_spmv_dev(h.h, dA.p, C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()))), stream)
        h.synchronize()
        y = dy.cpu().numpy()
        t0 = time.perf_counter()
        want = O.spmv(p[0], p[1], p[2], p[3], p[4], x)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        sabs = O.spmv(p[0], p[1], p[2], p[3], np.abs(p[4]), np.abs(x))
        ok = bool(np.all(np.abs(y - want) <= 1e-12 * sabs))
        by = G.algorithmic_bytes_spmv(p[0], p[1], len(p[3]), 8)
        print(json.dumps({"op": "spmv", "workload": "poisson2048 f64", "ms": ms, "gflops": 2 * len(p[3]) / ms / 1e6,
                          "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac_of_measured_peak": by / ms / 1e6 / peak,
                          "peak_source": src, "parity_ok": ok, "cpu_oracle_ms_1thread": cpu_ms}), flush=True)
        dA.free()

    if "ewise" in which:
        # C = A*A + A on the Poisson matrix (the add on either side of a product)
        p = G.poisson2d(2048)
        A = S.CsrMatrix(p[0], p[1], p[4], p[3], p[2])
        dA = S.DeviceCsr.upload(A, h)
        dB = dA.matmul(dA)
        outs = []

        def run_add():
            outs.append(dB.add(dA))
            if len(outs) > 1:
                outs.pop(0).free()
        ms = timed(run_add, stream, reps=10)
        got = outs[-1].download()
        b_host = dB.download()
        t0 = time.perf_counter()
        want = O.ewise((p[0], p[1], b_host.offsets, b_host.indices, b_host.vals), p, "add", True)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        ok = bool(np.array_equal(got.offsets, want[0]) and np.array_equal(got.indices, want[1]) and
                  np.array_equal(got.vals, want[2]))
        nb, na, nc = b_host.nnz(), len(p[3]), got.nnz()
        by = (na + nb + nc) * 12 + 3 * (p[0] + 1) * 8      # read A and B once, write C once (count pass not credited)
        print(json.dumps({"op": "ewise_add", "workload": "poisson2048 f64: A*A + A", "nnz_a": na, "nnz_b": nb, "nnz_c": nc,
                          "ms": ms, "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac_of_measured_peak": by / ms / 1e6 / peak,
                          "parity_ok": ok, "cpu_oracle_ms_1thread": cpu_ms}), flush=True)
        for d in outs:
            d.free()
        dB.free(); dA.free()

    if "dok" in which or "rect" in which:
        a = G.uniform_random(1_000_000, 4_000_000, 8, seed=5, dtype=np.int64, int_range=1 << 15)

    if "dok" in which:
        tr, tc, tv = G.triplets_with_rewrites(a, seed=5)
        n = len(tv)
        d_r = torch.from_numpy(tr.view(np.int64)).to(dev)
        d_c = torch.from_numpy(tc.view(np.int64)).to(dev)
        d_v = torch.from_numpy(tv).to(dev)
        torch.cuda.synchronize()
        outs = []

        def run():
            out = C.c_void_p()
            S._lib.check(h.h, L.spam_dok_to_csr_dev(h.h, 3, a[0], a[1], n, C.c_void_p(d_r.data_ptr()), C.c_void_p(d_c.data_ptr()),
                                                    C.c_void_p(d_v.data_ptr()), C.byref(out)))
            outs.append(out)
            if len(outs) > 1:
                L.spam_dcsr_free(h.h, outs.pop(0))
        ms = timed(run, stream, reps=10)
        got = S.DeviceCsr(h, outs[-1]).download()
        t0 = time.perf_counter()
        off, idx, val = O.dok_to_csr(a[0], a[1], tr, tc, tv)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        ok = bool(np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val))
        by = n * (8 + 8 + 8) + len(idx) * (4 + 8) + (a[0] + 1) * 8
        print(json.dumps({"op": "dok_to_csr", "workload": "C5 1Mx4M, 8/row, i64, shuffled stream with 1% rewrites, 0.1% deletes",
                          "triplets": n, "nnz_out": int(len(idx)), "ms": ms, "mtriplets_per_s": n / ms / 1e3,
                          "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac_of_measured_peak": by / ms / 1e6 / peak,
                          "parity_ok": ok, "cpu_oracle_ms_1thread": cpu_ms}), flush=True)

    if "rect" in which:
        A = S.CsrMatrix(a[0], a[1], a[4], a[3], a[2])
        dA = S.DeviceCsr.upload(A, h)
        # A^T on the device, checked bit-exact against the oracle's transpose
        ts = []

        def run_t():
            ts.append(dA.transpose())
            if len(ts) > 1:
                ts.pop(0).free()
        ms_t = timed(run_t, stream, reps=20)
        path_t = h.stats()["fallbacks"][4]
        dAT = ts[-1]
        gt = dAT.download()
        t0 = time.perf_counter()
        want_t = O.transpose(a)
        cpu_t = (time.perf_counter() - t0) * 1e3
        ok_t = bool(np.array_equal(gt.offsets, want_t[0]) and np.array_equal(gt.indices, want_t[1]) and
                    np.array_equal(gt.vals, want_t[2]))
        at = (a[1], a[0]) + want_t
        nnz_a = len(a[3])
        by_t = nnz_a * 12 * 2 + (a[0] + 1) * 8 + (a[1] + 1) * 8      # read A once, write A^T once
        print(json.dumps({"op": "transpose", "workload": "C5 1Mx4M, 8/row, i64", "nnz": nnz_a, "ms": ms_t,
                          "algorithmic_bytes": by_t, "gbs": by_t / ms_t / 1e6, "frac_of_measured_peak": by_t / ms_t / 1e6 / peak,
                          "parity_ok": ok_t, "path": {1: "counting", 2: "radix", 3: "bucket"}.get(path_t, path_t),
                          "cpu_oracle_ms_1thread": cpu_t}), flush=True)
        cs = []

        def run2():
            cs.append(dA.matmul(dAT))
            if len(cs) > 1:
                cs.pop(0).free()
        h.set_timing(True)
        ms = timed(run2, stream, reps=10)
        st = h.stats()
        h.set_timing(False)
        got = cs[-1].download()
        t0 = time.perf_counter()
        off, idx, val = O.mul_hash(a, at, True)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        ok = bool(np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val))
        by = G.algorithmic_bytes_spgemm(a[0], len(a[3]), st["flops"], st["nnz_c"], 8)
        print(json.dumps({"op": "spgemm", "workload": "C5 A*A^T, 1Mx4M, 8/row, i64 (bit-exact)", "products": st["flops"],
                          "nnz_c": st["nnz_c"], "ms": ms, "gflops": 2 * st["flops"] / ms / 1e6, "algorithmic_bytes": by,
                          "gbs": by / ms / 1e6, "frac_of_measured_peak": by / ms / 1e6 / peak, "parity_ok": ok,
                          "num_bin_rows": st["num_bin_rows"], "cpu_oracle_ms_all_threads": cpu_ms,
                          "cores": O.hardware_threads()}), flush=True)
    h.close()


if __name__ == "__main__":
    main()
