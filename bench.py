#!/usr/bin/env python
"""bench.py — SpGEMM throughput of the hot path (C = A*A, CsrMatrix::mul_hash) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl native|reference]

One "step" = one full product over the synthetic workload.  N=1 default workload = BASELINE.json
configs[1]: 2-D 5-point Poisson 2048x2048 grid, A*A, f64.  Prints ONE JSON line (rank 0).

  value      GFLOP/s (2 x intermediate products / step time), inputs resident in HBM, CUDA events
  e2e        same metric through the reference-facing two-phase C ABI with HOST (pinned) buffers:
             H2D of A (u64 indices), D2H of row_ptr/col_idx/val inside the timed region
  roofline   dominant kernel (numeric pass): algorithmic bytes / its event-timed duration
             against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the C++ restatement of spam_csr::mul_hash (oracle/, kind "port": the Rust reference
             cannot be built here) timed on this box's host cores
  --impl reference  times that CPU restatement alone, same metric/config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_DESC = {
    "poisson2048": "2-D 5-point Poisson 2048x2048 grid (4.19M rows, 20.96M nnz), A*A, f64  [BASELINE configs[1]]",
    "uniform10k": "uniform random CSR 10k x 10k, ~10 nnz/row, A*A, f64  [BASELINE configs[0]]",
    "stencil160": "3-D 27-point stencil 160^3 (4.1M rows, 109M nnz), A*A, f64  [BASELINE configs[2]]",
    "stencil96": "3-D 27-point stencil 96^3 (0.88M rows), A*A, f64  [reduced configs[2]]",
    "rmat22": "R-MAT(0.45,0.15,0.15,0.25) scale 22 ef 16, A*A, f64  [BASELINE configs[3]]",
    "rmat18": "R-MAT(0.45,0.15,0.15,0.25) scale 18 ef 16, A*A, f64  [reduced configs[3]]",
    "rmat20": "R-MAT(0.45,0.15,0.15,0.25) scale 20 ef 16, A*A, f64  [reduced configs[3]]",
}


def resolve_scaling(args):
    """`strong`: the named matrix, rows split over the ranks (BASELINE configs[2], [3]: "row-partitioned
    1/2/4/8 B200").  `weak`: every rank owns one unit of the named workload — for Poisson (configs[1], a
    1-GPU configuration) a 2048 x 2048 block of lines of a 2048 x (2048 N) grid, B = the whole matrix
    replicated; per-GPU work is fixed as N grows.  `auto` = weak for poisson2048, strong otherwise."""
    if args.scaling != "auto":
        if args.scaling == "weak" and args.workload != "poisson2048":
            raise SystemExit("--scaling weak is defined for poisson2048 only")
        return args.scaling
    return "weak" if (args.workload == "poisson2048" and args.gpus > 1) else "strong"


def make_workload(name, units=1):
    from sparse_matrix_b200 import generators as G
    if name == "poisson2048":
        return G.poisson2d(2048, ny=2048 * units)
    if name == "uniform10k":
        return G.uniform_random(10_000, 10_000, 10, seed=1)
    if name == "stencil160":
        return G.stencil27(160)
    if name == "stencil96":
        return G.stencil27(96)
    if name == "rmat22":
        return G.rmat(22)
    if name == "rmat18":
        return G.rmat(18)
    if name == "rmat20":
        return G.rmat(20)
    raise SystemExit(f"unknown workload {name}")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(mat, budget_s=20.0, max_runs=5):
    """Times the oracle (C++ restatement of mul_hash, all host threads, unsorted output like the
    reference bench `bench_mul::<false>`, spam_csr/src/lib.rs:403-410) on the same matrix."""
    from oracle import pyoracle as O
    O.build()
    a = (mat[0], mat[1], np.ascontiguousarray(mat[2], np.uint64), np.ascontiguousarray(mat[3], np.uint64), mat[4])
    best, runs, spent, flops = None, 0, 0.0, 0
    while runs < max_runs and (runs == 0 or spent + (best or 0) < budget_s):
        dt, nnz, flops = O.mul_hash_timed(a, a, False, 0)
        best = dt if best is None else min(best, dt)
        spent += dt
        runs += 1
    return {"value": 2.0 * flops / best / 1e9, "unit": "GFLOP/s", "cores": O.hardware_threads(), "kind": "port",
            "sample": f"full workload A*A, best of {runs} runs ({best * 1e3:.1f} ms each), unsorted output",
            "ms": best * 1e3}


def pinned_array(L, n, dtype, keep):
    p = C.c_void_p()
    nbytes = max(1, n) * np.dtype(dtype).itemsize
    st = L.spam_host_alloc(C.byref(p), nbytes)
    if st != 0:
        raise RuntimeError("spam_host_alloc failed")
    keep.append(p)
    buf = (C.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=n)


class DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_reference(args, rank, world):
    if rank != 0:
        return
    scaling = resolve_scaling(args)
    mat = make_workload(args.workload, args.gpus if scaling == "weak" else 1)
    from oracle import pyoracle as O
    O.build()
    a = (mat[0], mat[1], np.ascontiguousarray(mat[2], np.uint64), np.ascontiguousarray(mat[3], np.uint64), mat[4])
    for _ in range(args.warmup):
        O.mul_hash_timed(a, a, False, 0)
    t = []
    flops = 0
    for _ in range(args.steps):
        dt, nnz, flops = O.mul_hash_timed(a, a, False, 0)
        t.append(dt)
    ms = 1e3 * sum(t) / len(t)
    val = 2.0 * flops / (ms / 1e3) / 1e9
    line = {"impl": "reference", "metric": "spgemm_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": WORKLOAD_DESC[args.workload], "rows": int(mat[0])},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": O.hardware_threads(), "kind": "port",
                             "sample": "full workload A*A per step; C++ restatement of spam_csr::mul_hash "
                                       "(the Rust reference cannot be built: no cargo/rustc in the image)"},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="poisson2048", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import sparse_matrix_b200 as S
    from sparse_matrix_b200 import generators as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    handle = S.Handle(local_rank)
    # one explicit (non-default) stream for everything: the library's kernels, torch's events and the
    # NCCL collectives are all ordered on it, so the CUDA events below see the kernels they bracket
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    handle.set_stream(stream.cuda_stream)
    L = handle.L

    # ---- inputs: rank 0 generates, everyone gets A over NCCL (B = A is replicated) ----
    t_bcast = 0.0
    scaling = resolve_scaling(args)
    if world == 1 or rank == 0:
        mat = make_workload(args.workload, world if scaling == "weak" else 1)
        rows, cols = mat[0], mat[1]
        h_ptr = torch.from_numpy(np.ascontiguousarray(mat[2]).view(np.int64))
        h_idx = torch.from_numpy(np.ascontiguousarray(mat[3]).astype(np.uint32).view(np.int32))
        h_val = torch.from_numpy(np.ascontiguousarray(mat[4]))
        meta = torch.tensor([rows, cols, h_idx.shape[0]], dtype=torch.int64, device=dev)
    else:
        mat = None
        meta = torch.zeros(3, dtype=torch.int64, device=dev)
    if world > 1:
        from sparse_matrix_b200 import distributed as D
        dist.broadcast(meta, src=0)
    rows, cols, nnz_a = (int(x) for x in meta.tolist())
    if world == 1 or rank == 0:
        d_ptr, d_idx, d_val = h_ptr.to(dev), h_idx.to(dev), h_val.to(dev)
    else:
        d_ptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
        d_idx = torch.empty(nnz_a, dtype=torch.int32, device=dev)
        d_val = torch.empty(nnz_a, dtype=torch.float64, device=dev)
    if world > 1:
        t_bcast = D.replicate([d_ptr, d_idx, d_val], src=0)
    dA = S.DeviceCsr.wrap(handle, np.float64, rows, cols, nnz_a, d_ptr.data_ptr(), d_idx.data_ptr(), d_val.data_ptr(),
                          keepalive=(d_ptr, d_idx, d_val))

    # ---- one step ----
    gathered_ms = None
    if world == 1:
        def step():
            c = dA.matmul(dA)
            c.free()
    else:
        # Distributed input layout = A row-sharded by flop-balanced blocks (the rows_to_threads formula
        # with tnum = world, mul_hash.rs:51-62), B replicated.  Building that layout is set-up, like the
        # reference building its CsrMatrix outside the timed closure (lib.rs:403-410); it is timed and
        # reported as partition_ms.  Each rank's product still does its own flop count / binning per step.
        torch.cuda.synchronize()
        tp0 = time.perf_counter()
        starts, total = dA.rows_to_parts(dA, world, balance="cost")
        blk = dA.slice_rows(int(starts[rank]), int(starts[rank + 1]))
        handle.synchronize()
        partition_ms = (time.perf_counter() - tp0) * 1e3

        def step(gather=False):
            c = blk.matmul(dA)
            if gather:
                i = c.info()
                lp = torch.as_tensor(DevArray(i["d_ptr"], i["rows"] + 1, "<i8"), device=dev)
                li = torch.as_tensor(DevArray(i["d_idx"], max(1, i["nnz"]), "<i4"), device=dev)[:i["nnz"]]
                lv = torch.as_tensor(DevArray(i["d_val"], max(1, i["nnz"]), "<f8"), device=dev)[:i["nnz"]]
                rows_per = [int(starts[r + 1] - starts[r]) for r in range(world)]
                D.gathered_csr(lp, li, lv, rows_per)
            c.free()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks are sampled from before the warm-up to the end of the timed region; the warm-up is
    # stretched to >= 0.4 s of the same load so that nvidia-smi (20 ms period) sees the steady state
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_w = time.perf_counter()
    n_warm = 0
    while n_warm < args.warmup or (time.perf_counter() - t_w < 0.4 and n_warm < 2000):
        step()
        n_warm += 1
    if world > 1:  # same count on every rank (the gather variant below is collective)
        tw = torch.tensor([n_warm], dtype=torch.int64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    sync_all()
    args.warmup = n_warm

    handle.set_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"ms_flop": 0.0, "ms_symbolic": 0.0, "ms_scan": 0.0, "ms_numeric": 0.0, "ms_total": 0.0}
    launches = 0
    sync_all()
    ev0.record()
    handle.phase_totals(reset=True)   # the library sums each product's phase events; nothing is read back in the loop
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    st = handle.stats()               # last product: counts and launches (identical every step)
    tot = handle.phase_totals()
    assert tot["products"] == args.steps, tot
    for k in phase:
        phase[k] = tot[k]
    launches = st["kernel_launches"] * args.steps
    handle.set_timing(False)
    ms_step = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(per_rank, t)
        rank_ms = [round(float(x.item()), 4) for x in per_rank]
        ms_step = max(rank_ms)
    flops, nnz_c = st["flops"], st["nnz_c"]
    local_nnz_c = nnz_c
    if world > 1:
        t = torch.tensor([flops, nnz_c, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        flops, nnz_c, launches = (int(x) for x in t.tolist())
        # the same step including the all-gather-v of C (reported separately, SURVEY §7 hard parts)
        for _ in range(2):
            step(gather=True)
        sync_all()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        nrep = max(1, min(args.steps, 5))
        for _ in range(nrep):
            step(gather=True)
        g1.record()
        sync_all()
        t = torch.tensor([g0.elapsed_time(g1) / nrep], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered_ms = float(t.item())

    val_size = 8
    bytes_alg = G.algorithmic_bytes_spgemm(rows, nnz_a, flops, nnz_c, val_size)
    peak, peak_src = measured_peak()
    gflops = 2.0 * flops / (ms_step / 1e3) / 1e9

    line = {"metric": "spgemm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": WORKLOAD_DESC[args.workload], "rows": rows, "nnz_a": nnz_a,
                       "products": flops, "nnz_c": nnz_c, "algorithmic_bytes": bytes_alg,
                       "l2": "no flush: per-step working set (A + C, %.2f GB) exceeds the 126 MB L2" %
                             ((nnz_a * 12 + nnz_c * 12 + rows * 16) / 1e9),
                       "sharding": "single GPU" if world == 1 else
                                   f"A pre-sharded in device-cost-balanced row blocks (spam_rows_to_parts_cost) over {world} ranks (partition + slice "
                                   f"{partition_ms:.2f} ms, untimed set-up), B replicated (NCCL broadcast "
                                   f"{t_bcast * 1e3:.1f} ms, untimed), C left row-sharded in `value`; `gathered` adds "
                                   f"the all-gather-v"},
            "gpu_launches": launches,
            "hbm_gbs_pipeline": bytes_alg / (ms_step / 1e3) / 1e9,
            "phases_ms": {k: v / args.steps for k, v in phase.items()}}
    if world > 1:
        line["rank_ms"] = rank_ms      # each rank's own device time per step: the load balance of the partition
        if scaling == "weak":
            line["config"]["weak_unit"] = ("one 2048 x 2048 block of grid lines (4.19M rows of A) per GPU; the matrix is "
                                           f"the Poisson operator on a 2048 x {2048 * world} grid, B = all of it, replicated")
    if gathered_ms is not None:
        line["gathered"] = {"ms_per_step": gathered_ms, "value": 2.0 * flops / (gathered_ms / 1e3) / 1e9,
                            "unit": "GFLOP/s", "note": "same step plus all-gather-v of row_ptr/col_idx/val so every "
                                                       "rank holds the full C"}

    if rank == 0:
        # ---- roofline of the dominant kernel: the numeric pass (one k_num_* launch per non-empty bin) ----
        ms_num = phase["ms_numeric"] / args.steps
        my_bytes = bytes_alg if world == 1 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload, {}).get("numeric_dram_bytes")
            except Exception:
                traffic = None
        if my_bytes is not None and ms_num > 0:
            ach = my_bytes / (ms_num / 1e3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": traffic,
                                "kernel": "numeric pass (k_num_merge<double,6,128> on Poisson: one launch per step)",
                                "kernel_ms": ms_num, "peak_source": peak_src,
                                "pipeline_frac": line["hbm_gbs_pipeline"] / peak}
        else:
            ach = line["hbm_gbs_pipeline"]
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s",
                                "frac": ach / (peak * world), "traffic": None, "kernel": "whole sharded step",
                                "peak_source": peak_src + f" x {world} GPUs"}
        line["clocks"] = clocks

    # ---- e2e: the reference-facing two-phase C ABI with pinned HOST buffers.  At N > 1 every rank is a
    # host caller multiplying ITS row block of A (host memory) by all of B (host memory): both are uploaded
    # and the C shard is downloaded inside the timed region; time = max over ranks.
    do_e2e = not args.no_e2e
    if do_e2e and world > 1:
        # bound pinned host memory: skip (and say so) when the box cannot hold every rank's buffers
        need = world * ((rows + 1) * 8 + nnz_a * 16) + (rows + 1) * 16 + nnz_a * 16 + nnz_c * 16
        try:
            avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
        except Exception:
            avail = 0
        flag = torch.tensor([1 if need * 3 > avail else 0], dtype=torch.int64, device=dev)
        dist.broadcast(flag, src=0)     # one decision for every rank (the e2e steps are bracketed by barriers)
        if int(flag.item()):
            do_e2e = False
            if rank == 0:
                line["e2e_skipped"] = f"pinned host buffers for {world} ranks need {need / 1e9:.1f} GB, MemAvailable {avail / 1e9:.1f} GB"
    if do_e2e:
        keep = []
        try:
            if world == 1:
                hb_ptr, hb_idx, hb_val = mat[2], mat[3], mat[4]
                a_rows, a_nnz = rows, nnz_a
            else:
                hb_ptr = d_ptr.cpu().numpy().view(np.uint64)
                hb_idx = d_idx.cpu().numpy().view(np.uint32)
                hb_val = d_val.cpu().numpy()
                r0, r1 = int(starts[rank]), int(starts[rank + 1])
                a_rows = r1 - r0
                e0_, e1_ = int(hb_ptr[r0]), int(hb_ptr[r1])
                a_nnz = e1_ - e0_
            p_ptr = pinned_array(L, rows + 1, np.uint64, keep); p_ptr[:] = hb_ptr
            p_idx = pinned_array(L, nnz_a, np.uint64, keep); p_idx[:] = hb_idx
            p_val = pinned_array(L, nnz_a, np.float64, keep); p_val[:] = hb_val
            if world == 1:
                pa_ptr, pa_idx, pa_val = p_ptr, p_idx, p_val     # A aliases B: the library uploads it once
            else:
                pa_ptr = pinned_array(L, a_rows + 1, np.uint64, keep); pa_ptr[:] = hb_ptr[r0:r1 + 1] - hb_ptr[r0]
                pa_idx = pinned_array(L, a_nnz, np.uint64, keep); pa_idx[:] = hb_idx[e0_:e1_]
                pa_val = pinned_array(L, a_nnz, np.float64, keep); pa_val[:] = hb_val[e0_:e1_]
            c_ptr = pinned_array(L, a_rows + 1, np.uint64, keep)
            c_idx = pinned_array(L, local_nnz_c, np.uint64, keep)
            c_val = pinned_array(L, local_nnz_c, np.float64, keep)

            def e2e_step():
                nz = C.c_uint64()
                S._lib.check(handle.h, L.spam_spgemm_symbolic(handle.h, 1, a_rows, cols, S._lib.ptr(pa_ptr),
                                                              S._lib.ptr(pa_idx), S._lib.ptr(pa_val), rows, cols,
                                                              S._lib.ptr(p_ptr), S._lib.ptr(p_idx), S._lib.ptr(p_val),
                                                              S._lib.ptr(c_ptr), C.byref(nz)))
                assert nz.value == local_nnz_c
                S._lib.check(handle.h, L.spam_spgemm_numeric(handle.h, S._lib.ptr(c_idx), S._lib.ptr(c_val), 1))

            for _ in range(2):
                e2e_step()
            sync_all()
            k = max(3, min(args.steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                e2e_step()
            e1.record()
            sync_all()
            ms_e2e = e0.elapsed_time(e1) / k
            # bytes actually moved over PCIe by the last step, as counted by the library (at N > 1 only the band
            # of B that this rank's rows of A reference is uploaded)
            st_e2e = handle.stats()
            h2d, d2h = int(st_e2e["bytes_h2d"]), int(st_e2e["bytes_d2h"])
            if world > 1:
                t = torch.tensor([ms_e2e, -ms_e2e, h2d, d2h], dtype=torch.float64, device=dev)
                tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                ms_e2e, h2d, d2h = float(tmax[0].item()), int(t[2].item()), int(t[3].item())
            line["e2e"] = {"value": 2.0 * flops / (ms_e2e / 1e3) / 1e9, "unit": "GFLOP/s",
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                           "api": "spam_spgemm_symbolic + spam_spgemm_numeric (host u64 indices, pinned buffers; " +
                                  ("A aliases B so it is uploaded once)" if world == 1 else
                                   "every rank uploads its row block of A and the rows of B it references, downloads its shard of C; "
                                   "bytes summed over ranks, time = max over ranks)")}
        finally:
            for p in keep:
                L.spam_host_free(p)
    elif rank == 0:
        line["e2e"] = None

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(mat)
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        blk.free()
    dA.free()
    handle.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
