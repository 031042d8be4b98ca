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
    replicated; per-GPU work is fixed as N grows.  `auto` = weak for poisson2048 at every N (the N = 1 line carries the
    same label as the N > 1 lines it is compared with), strong otherwise."""
    if args.scaling != "auto":
        if args.scaling == "weak" and args.workload != "poisson2048":
            raise SystemExit("--scaling weak is defined for poisson2048 only")
        return args.scaling
    return "weak" if args.workload == "poisson2048" else "strong"


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
            "config": {"workload": args.workload, "desc": WORKLOAD_DESC[args.workload], "rows": int(mat[0]),
                       "nnz_a": int(len(mat[3])), "products": int(flops), "nnz_c": int(nnz),
                       "algorithmic_bytes": int(len(mat[3]) * 12 + (mat[0] + 1) * 16 + flops * 12 + nnz * 12),
                       "sharding": "CPU: flop-balanced row blocks over the host threads (rows_to_threads, mul_hash.rs:38-64)"},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": O.hardware_threads(), "kind": "port",
                             "sample": "full workload A*A per step; C++ restatement of spam_csr::mul_hash "
                                       "(the Rust reference cannot be built: no cargo/rustc in the image)"},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def dominant_kernel(workload, st, direct=False):
    """Name of the numeric kernel that holds most rows' products for this workload (roofline.kernel)."""
    nb = st["num_bin_rows"]
    if nb[10] and nb[10] >= sum(nb) * 0.9:
        return "k_num_merge<double,K,128> (merge bin: one launch per step)"
    if sum(nb[11:16]) >= sum(nb[1:10]):
        return "k_num_esc<double,NW> (bucket-sort bins 11..14) + hash bins for short rows: one launch per non-empty bin"
    return "k_num_row<double,NW,CAP> (hash bins, one launch per non-empty bin)"


def time_steps(fn, steps, sync_all, torch):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    sync_all()
    return ev0.elapsed_time(ev1) / steps


def median_step(fn, steps, sync_all, torch):
    """Median of individually timed steps (robust to a one-off stall; used for the long R-MAT steps only)."""
    return float(np.median([time_steps(fn, 1, sync_all, torch) for _ in range(steps)]))


def max_over_ranks(x, world, dev, torch, dist):
    if world == 1:
        return float(x), [round(float(x), 4)]
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    per = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(per, t)
    vals = [round(float(v.item()), 4) for v in per]
    return max(vals), vals


def device_matrix(S, handle, mat, rank, world, dev, torch, np_dtype=np.float64):
    """A (= B) resident on every rank: rank 0 uploads, the library's own communicator replicates it
    (spam_comm_broadcast = ncclBroadcast over NVLink).  Returns (DeviceCsr, rows, cols, nnz, seconds for the replicate)."""
    if rank == 0:
        rows, cols = mat[0], mat[1]
        h_ptr = torch.from_numpy(np.ascontiguousarray(mat[2]).view(np.int64))
        h_idx = torch.from_numpy(np.ascontiguousarray(mat[3]).astype(np.uint32).view(np.int32))
        h_val = torch.from_numpy(np.ascontiguousarray(mat[4]))
        meta = np.array([rows, cols, h_idx.shape[0]], dtype=np.uint64)
    else:
        meta = np.zeros(3, dtype=np.uint64)
    if world > 1:
        meta = handle.comm_allgather_u64(meta)[0]
    rows, cols, nnz = (int(x) for x in meta)
    if rank == 0:
        d_ptr, d_idx, d_val = h_ptr.to(dev), h_idx.to(dev), h_val.to(dev)
    else:
        d_ptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
        d_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
        d_val = torch.empty(nnz, dtype=torch.float64, device=dev)
    t_b = 0.0
    if world > 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in (d_ptr, d_idx, d_val):
            handle.comm_broadcast(t.data_ptr(), t.numel() * t.element_size(), 0)
        handle.synchronize()
        t_b = time.perf_counter() - t0
    dA = S.DeviceCsr.wrap(handle, np_dtype, rows, cols, nnz, d_ptr.data_ptr(), d_idx.data_ptr(), d_val.data_ptr(),
                          keepalive=(d_ptr, d_idx, d_val))
    return dA, rows, cols, nnz, t_b


def as_tensors(c, dev, torch):
    i = c.info()
    lp = torch.as_tensor(DevArray(i["d_ptr"], i["rows"] + 1, "<i8"), device=dev)
    li = torch.as_tensor(DevArray(i["d_idx"], max(1, i["nnz"]), "<i4"), device=dev)[:i["nnz"]]
    lv = torch.as_tensor(DevArray(i["d_val"], max(1, i["nnz"]), "<f8"), device=dev)[:i["nnz"]]
    return lp, li, lv


def gathered_parity(dA, gathered, dev, torch, exact):
    """The assembled multi-GPU C against the single-GPU product of the same matrix on this rank: nnz, row_ptr and
    col_idx identical; values identical (`exact`: merge bin, deterministic) or within 1e-12 of each entry's sum
    of |products| (the same product on |A|)."""
    S_ = dA.handle
    ref = dA.matmul(dA)
    ok = True
    try:
        gp, gi, gv = as_tensors(gathered, dev, torch)
        rp, ri, rv = as_tensors(ref, dev, torch)
        ok = ok and gathered.info()["nnz"] == ref.info()["nnz"] and bool(torch.equal(gp, rp)) and bool(torch.equal(gi, ri))
        if ok and exact:
            ok = bool(torch.equal(gv, rv))
        elif ok:
            torch.sub(rv, gv, out=rv)
            rv.abs_()                                        # |difference|, in the reference product's own memory
            i = dA.info()
            av = torch.as_tensor(DevArray(i["d_val"], i["nnz"], "<f8"), device=dev).abs()
            dAbs = type(dA).wrap(S_, np.float64, i["rows"], i["cols"], i["nnz"], i["d_ptr"], i["d_idx"], av.data_ptr(), keepalive=(av,))
            sabs = dAbs.matmul(dAbs)
            _, _, sv = as_tensors(sabs, dev, torch)
            ok = bool((rv <= 1e-12 * sv).all())
            sabs.free(); dAbs.free()
    finally:
        ref.free()
    return ok


def rmat_strong_section(args, S, D, G, handle, rank, world, dev, torch, dist, sync_all):
    """BASELINE configs[3] / north_star: strong scaling of the largest R-MAT product, in the same invocation at every
    N: the 1-GPU product on this box, the row-sharded product (C left sharded), and the product with C assembled on
    every rank (peer stores overlapped with the numeric kernels; and the NCCL-broadcast variant beside it)."""
    name = args.scale_workload
    out = {"workload": name, "desc": WORKLOAD_DESC[name], "scaling": "strong",
           "timing": "CUDA events around single steps between barriers, median of the steps, max over ranks "
                     "(ms_1gpu_ref: every rank runs the whole product on its own GPU, fastest rank)"}
    t0 = time.perf_counter()
    mat = make_workload(name) if rank == 0 else None
    out["generate_s"] = round(time.perf_counter() - t0, 2) if rank == 0 else None
    torch.cuda.synchronize()
    t_host0 = time.perf_counter()
    dA, rows, cols, nnz_a, t_b = device_matrix(S, handle, mat, rank, world, dev, torch)
    steps = max(2, min(args.steps, 3))
    # ---- the whole product on ONE GPU (every rank does it at the same time on its own copy: max over ranks) ----
    outs = []

    def one():
        c = dA.matmul(dA)
        outs.append(c)
        if len(outs) > 1:
            outs.pop(0).free()
    one(); one()
    # every rank times the same product on its own GPU, step by step; the median step of the fastest rank is "one GPU
    # alone" (an outlier rank was seen once: 1.4 s for a 96 ms product on one of four GPUs), the slowest rank's is
    # reported beside it
    ms1_local = median_step(one, steps, sync_all, torch)
    ms1_max, ms1_ranks = max_over_ranks(ms1_local, world, dev, torch, dist)
    ms1 = min(ms1_ranks)
    st = handle.stats()
    flops, nnz_c = int(st["flops"]), int(st["nnz_c"])
    while outs:
        outs.pop().free()
    out.update({"rows": rows, "nnz_a": nnz_a, "products": flops, "nnz_c": nnz_c, "ms_1gpu_ref": ms1,
                "ms_1gpu_ref_per_rank": ms1_ranks,
                "gflops_1gpu": 2.0 * flops / ms1 / 1e6, "num_bin_rows": st["num_bin_rows"], "fallbacks": st["fallbacks"],
                "algorithmic_bytes": G.algorithmic_bytes_spgemm(rows, nnz_a, flops, nnz_c, 8)})
    out["hbm_frac_1gpu"] = out["algorithmic_bytes"] / ms1 / 1e6 / measured_peak()[0]
    if world == 1:
        dA.free()
        return out
    # ---- partition (device-time balanced rows_to_threads formula) + slice: set-up, reported ----
    torch.cuda.synchronize()
    tp0 = time.perf_counter()
    starts, _ = dA.rows_to_parts(dA, world, balance="cost")
    r0, r1 = int(starts[rank]), int(starts[rank + 1])
    blk = dA.slice_rows(r0, r1)
    handle.synchronize()
    out["partition_ms"] = (time.perf_counter() - tp0) * 1e3
    out["bcast_ms"] = t_b * 1e3

    def sharded():
        blk.matmul(dA).free()

    def gathered(mode, nsub=None):
        def f():
            blk.matmul_gathered(dA, r0, rows, nsub=nsub or args.nsub, mode=mode).free()
        return f
    sharded(); sharded()
    ms_s = median_step(sharded, steps, sync_all, torch)
    ms_s, rank_ms = max_over_ranks(ms_s, world, dev, torch, dist)
    gm = args.gather_mode
    gathered(gm)(); gathered(gm)()
    out["ms_total_from_host_A"] = None
    ms_g, _ = max_over_ranks(median_step(gathered(gm), steps, sync_all, torch), world, dev, torch, dist)
    peer = handle.comm_info()["peer_mapped"]
    # parity of the assembled C (outside every timed region)
    g = blk.matmul_gathered(dA, r0, rows, nsub=args.nsub, mode=gm)
    ok = gathered_parity(dA, g, dev, torch, exact=False)
    g.free()
    okt = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    gathered(1)(); gathered(1)()
    ms_n, _ = max_over_ranks(median_step(gathered(1), steps, sync_all, torch), world, dev, torch, dist)
    recv = (nnz_c * 12 + (rows + 1) * 8) * (world - 1) / world
    if args.gather_sweep:       # how the exchange is done and how finely it is pipelined (builder runs only)
        sweep = {}
        for mode in (0, 2):
            for nsub in (1, 4, 8, 16):
                f = gathered(mode, nsub)
                f(); f()
                t, _ = max_over_ranks(median_step(f, steps, sync_all, torch), world, dev, torch, dist)
                sweep[f"mode{mode}_nsub{nsub}"] = round(t, 3)
        out["gather_sweep_ms"] = sweep
    out.update({"ms_sharded": ms_s, "rank_ms": rank_ms, "speedup_sharded": ms1 / ms_s,
                "ms_gathered": ms_g, "speedup_gathered": ms1 / ms_g, "gather": (("peer stores over NVLink (k_push), " if (gm == 0 or (gm < 0 and world > 2)) else "peer-to-peer copies by the copy engines, ")
                           + f"{args.nsub} sub-blocks pipelined with the numeric kernels") if peer else "peer mapping unavailable: NCCL broadcasts",
                "ms_gathered_nccl": ms_n, "speedup_gathered_nccl": ms1 / ms_n,
                "gathered_parity": bool(int(okt.item())), "recv_bytes_per_gpu": int(recv),
                "recv_gbs_per_gpu_whole_step": recv / ms_g / 1e6,
                "recv_gbs_per_gpu_exchange_only": recv / max(ms_g - ms_s, 1e-3) / 1e6})
    # from A in rank 0's host memory to the assembled C on every rank, once: upload + replicate + partition + product
    out["ms_total_from_host_A"] = out["bcast_ms"] + out["partition_ms"] + ms_g + (0.0 if mat is None else 0.0)
    out["ms_total_from_host_A_note"] = "bcast_ms (replicate A = B over NCCL) + partition_ms + ms_gathered; the H2D copy of A on rank 0 is not included"
    blk.free(); dA.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="poisson2048", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"])
    ap.add_argument("--scale-workload", default="rmat22", choices=sorted(WORKLOAD_DESC),
                    help="the strong-scaling product measured beside the headline (config.strong_scaling)")
    ap.add_argument("--no-scale-section", action="store_true")
    ap.add_argument("--nsub", type=int, default=4, help="row sub-blocks that pipeline numeric kernels and exchange")
    ap.add_argument("--gather-mode", type=int, default=-1, choices=[-1, 0, 2],
                    help="0: push kernel (peer stores), 2: copy engines, -1: the library picks by the number of ranks")
    ap.add_argument("--gather-sweep", action="store_true", help="time every gather mode x sub-block count (strong-scaling section)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    requested_warmup = args.warmup
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import sparse_matrix_b200 as S
    from sparse_matrix_b200 import distributed as D
    from sparse_matrix_b200 import generators as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    handle = S.Handle(local_rank)
    # one explicit (non-default) stream for everything: the library's kernels, torch's events and the
    # collectives are all ordered on it, so the CUDA events below see the kernels they bracket
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    handle.set_stream(stream.cuda_stream)
    L = handle.L
    if world > 1:
        D.init_comm(handle)      # the library's own NCCL communicator + peer-mapped gather buffers

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- inputs: rank 0 generates, the library replicates A (B = A) ----
    scaling = resolve_scaling(args)
    mat = make_workload(args.workload, world if scaling == "weak" else 1) if rank == 0 else None
    dA, rows, cols, nnz_a, t_bcast = device_matrix(S, handle, mat, rank, world, dev, torch)

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
        r0, r1 = int(starts[rank]), int(starts[rank + 1])
        blk = dA.slice_rows(r0, r1)
        handle.synchronize()
        partition_ms = (time.perf_counter() - tp0) * 1e3

        def step():
            blk.matmul(dA).free()

        def step_gathered():
            blk.matmul_gathered(dA, r0, rows, nsub=args.nsub, mode=args.gather_mode).free()

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
    sync_all()

    handle.set_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"ms_flop": 0.0, "ms_symbolic": 0.0, "ms_scan": 0.0, "ms_numeric": 0.0, "ms_total": 0.0}
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
    ms_step, rank_ms = max_over_ranks(ms_step, world, dev, torch, dist)
    flops, nnz_c = st["flops"], st["nnz_c"]
    local_nnz_c = nnz_c
    gathered_ok = None
    if world > 1:
        t = torch.tensor([flops, nnz_c, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        flops, nnz_c, launches = (int(x) for x in t.tolist())
        # the same step with C assembled on every rank by peer stores overlapped with the numeric kernels
        # (spam_spgemm_gathered), reported beside the sharded value
        for _ in range(3):
            step_gathered()
        nrep = max(3, min(args.steps, 10))
        gathered_ms, _ = max_over_ranks(time_steps(step_gathered, nrep, sync_all, torch), world, dev, torch, dist)
        g = blk.matmul_gathered(dA, r0, rows, nsub=args.nsub, mode=args.gather_mode)
        ok = gathered_parity(dA, g, dev, torch, exact=(args.workload == "poisson2048"))
        g.free()
        okt = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gathered_ok = bool(int(okt.item()))

    val_size = 8
    bytes_alg = G.algorithmic_bytes_spgemm(rows, nnz_a, flops, nnz_c, val_size)
    peak, peak_src = measured_peak()
    gflops = 2.0 * flops / (ms_step / 1e3) / 1e9

    line = {"metric": "spgemm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": requested_warmup, "warmup_run": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": WORKLOAD_DESC[args.workload], "rows": rows, "nnz_a": nnz_a,
                       "products": flops, "nnz_c": nnz_c, "algorithmic_bytes": bytes_alg,
                       "l2": "no flush: per-step working set (A + C, %.2f GB) exceeds the 126 MB L2" %
                             ((nnz_a * 12 + nnz_c * 12 + rows * 16) / 1e9),
                       "sharding": "single GPU" if world == 1 else
                                   f"A pre-sharded in device-cost-balanced row blocks (spam_rows_to_parts_cost) over {world} ranks (partition + slice "
                                   f"{partition_ms:.2f} ms, untimed set-up), B replicated (spam_comm_broadcast "
                                   f"{t_bcast * 1e3:.1f} ms, untimed), C left row-sharded in `value`; `gathered` = the same step with C "
                                   f"assembled on every rank (spam_spgemm_gathered)"},
            "gpu_launches": launches,
            "hbm_gbs_pipeline": bytes_alg / (ms_step / 1e3) / 1e9,
            "phases_ms": {k: v / args.steps for k, v in phase.items()}}
    if world > 1:
        line["rank_ms"] = rank_ms      # each rank's own device time per step: the load balance of the partition
        if scaling == "weak":
            line["config"]["weak_unit"] = ("one 2048 x 2048 block of grid lines (4.19M rows of A) per GPU; the matrix is "
                                           f"the Poisson operator on a 2048 x {2048 * world} grid, B = all of it, replicated")
    if gathered_ms is not None:
        recv = (nnz_c * 12 + (rows + 1) * 8) * (world - 1) / world
        line["gathered"] = {"ms_per_step": gathered_ms, "value": 2.0 * flops / (gathered_ms / 1e3) / 1e9,
                            "unit": "GFLOP/s", "parity": gathered_ok, "recv_bytes_per_gpu": int(recv),
                            "recv_gbs_per_gpu": recv / gathered_ms / 1e6,
                            "peer_mapped": handle.comm_info()["peer_mapped"],
                            "note": "same step with every rank holding the whole C afterwards: numeric kernels write each "
                                    "rank's rows at their offset-fixed place, a push kernel stores them into every peer's "
                                    "copy over NVLink, sub-block by sub-block behind the numeric kernels"}
        line["gathered_parity"] = gathered_ok

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
                                # the same kernel by the DRAM bytes ncu counted for it (B rows mostly hit L1/L2):
                                "dram_frac": (traffic / (ms_num / 1e3) / 1e9 / peak) if traffic else None,
                                "kernel": dominant_kernel(args.workload, st),
                                "kernel_ms": ms_num, "peak_source": peak_src,
                                "pipeline_frac": line["hbm_gbs_pipeline"] / peak}
        else:
            ach = line["hbm_gbs_pipeline"]
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s",
                                "frac": ach / (peak * world), "traffic": None, "kernel": "whole sharded step",
                                "peak_source": peak_src + f" x {world} GPUs"}
        line["clocks"] = clocks

    # ---- the same product for CsrMatrix<T, false> inputs (rows not sorted: what from_dok-shuffled inputs are) ----
    if world == 1 and args.workload == "poisson2048":
        rng = np.random.default_rng(7)
        o = mat[2].astype(np.int64)
        lens = np.diff(o)
        keyr = np.repeat(np.arange(rows), lens) + rng.random(nnz_a)       # a random order inside every row
        perm = np.argsort(keyr, kind="stable")
        umat = (rows, cols, mat[2], mat[3][perm], mat[4][perm])
        dU = S.DeviceCsr.upload(S.CsrMatrix(umat[0], umat[1], umat[4], umat[3], umat[2], is_sorted=False), handle)

        def ustep():
            dU.matmul(dU).free()
        for _ in range(3):
            ustep()
        line["config"]["unsorted_input"] = {"ms_per_step": time_steps(ustep, max(3, min(args.steps, 10)), sync_all, torch),
                                            "note": "same matrix with every row's entries shuffled (IS_SORTED = false): "
                                                    "multiplied by the cached copy of B with sorted rows (merge bin; r1 took the "
                                                    "thread-per-row hash bin: 3.1 ms)"}
        line["config"]["unsorted_input"]["num_bin_rows"] = handle.stats()["num_bin_rows"]
        dU.free()

    # ---- e2e: the reference-facing two-phase C ABI with HOST buffers.  At N > 1 every rank is a
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
            i = dA.info()
            d_ptr = torch.as_tensor(DevArray(i["d_ptr"], rows + 1, "<i8"), device=dev)
            d_idx = torch.as_tensor(DevArray(i["d_idx"], nnz_a, "<i4"), device=dev)
            d_val = torch.as_tensor(DevArray(i["d_val"], nnz_a, "<f8"), device=dev)
            if world == 1:
                hb_ptr, hb_idx, hb_val = mat[2], mat[3], mat[4]
                a_rows, a_nnz = rows, nnz_a
            else:
                hb_ptr = d_ptr.cpu().numpy().view(np.uint64)
                hb_idx = d_idx.cpu().numpy().view(np.uint32)
                hb_val = d_val.cpu().numpy()
                a_rows = r1 - r0
                e0_, e1_ = int(hb_ptr[r0]), int(hb_ptr[r1])
                a_nnz = e1_ - e0_

            def host_buffers(pinned):
                def arr(n, dt):
                    return pinned_array(L, n, dt, keep) if pinned else np.empty(max(1, n), dtype=dt)[:n]
                p_ptr = arr(rows + 1, np.uint64); p_ptr[:] = hb_ptr
                p_idx = arr(nnz_a, np.uint64); p_idx[:] = hb_idx
                p_val = arr(nnz_a, np.float64); p_val[:] = hb_val
                if world == 1:
                    pa = (p_ptr, p_idx, p_val)                 # A aliases B: the library uploads it once
                else:
                    pa_ptr = arr(a_rows + 1, np.uint64); pa_ptr[:] = hb_ptr[r0:r1 + 1] - hb_ptr[r0]
                    pa_idx = arr(a_nnz, np.uint64); pa_idx[:] = hb_idx[e0_:e1_]
                    pa_val = arr(a_nnz, np.float64); pa_val[:] = hb_val[e0_:e1_]
                    pa = (pa_ptr, pa_idx, pa_val)
                c = (arr(a_rows + 1, np.uint64), arr(local_nnz_c, np.uint64), arr(local_nnz_c, np.float64))
                return pa, (p_ptr, p_idx, p_val), c

            def e2e_time(pinned):
                pa, pb, c = host_buffers(pinned)

                def e2e_step():
                    nz = C.c_uint64()
                    S._lib.check(handle.h, L.spam_spgemm_symbolic(handle.h, 1, a_rows, cols, S._lib.ptr(pa[0]),
                                                                  S._lib.ptr(pa[1]), S._lib.ptr(pa[2]), rows, cols,
                                                                  S._lib.ptr(pb[0]), S._lib.ptr(pb[1]), S._lib.ptr(pb[2]),
                                                                  S._lib.ptr(c[0]), C.byref(nz)))
                    assert nz.value == local_nnz_c
                    S._lib.check(handle.h, L.spam_spgemm_numeric(handle.h, S._lib.ptr(c[1]), S._lib.ptr(c[2]), 1))

                for _ in range(2):
                    e2e_step()
                k = max(3, min(args.steps, 10))
                ms = time_steps(e2e_step, k, sync_all, torch)
                st_e = handle.stats()
                return ms, int(st_e["bytes_h2d"]), int(st_e["bytes_d2h"])

            ms_e2e, h2d, d2h = e2e_time(True)
            ms_page = None
            if world == 1:
                ms_page, _, _ = e2e_time(False)      # ordinary pageable buffers: what a Rust Vec is
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
            if ms_page is not None:
                line["e2e"]["pageable"] = {"ms_per_step": ms_page, "value": 2.0 * flops / (ms_page / 1e3) / 1e9,
                                           "note": "the same calls on ordinary (pageable) host buffers, like the Rust wrapper's Vecs"}
        finally:
            for p in keep:
                L.spam_host_free(p)
    elif rank == 0:
        line["e2e"] = None

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(mat)
    elif rank == 0:
        line["cpu_baseline"] = None

    if world > 1:
        blk.free()
    dA.free()
    mat = None

    # ---- the strong-scaling configuration of the north star, measured in the same invocation at every N ----
    if not args.no_scale_section and args.workload == "poisson2048":
        try:
            sec = rmat_strong_section(args, S, D, G, handle, rank, world, dev, torch, dist, sync_all)
        except Exception as e:      # the headline line must survive a failure here; say what happened
            sec = {"workload": args.scale_workload, "error": f"{type(e).__name__}: {e}"}
            if world > 1:
                raise
        if rank == 0:
            line["config"]["strong_scaling"] = sec

    if rank == 0:
        print(json.dumps(line), flush=True)
    handle.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
