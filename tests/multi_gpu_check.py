#!/usr/bin/env python
"""Multi-rank check of the in-library exchange (run under torchrun on >= 2 GPUs of one node; not a pytest file:
`pytest -m gpu` runs on one GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multi_gpu_check.py

Every rank: spam_comm_init, A replicated with spam_comm_broadcast, flop-balanced row blocks, spam_spgemm_gathered in
both gather modes (peer stores / grouped ncclBroadcast) and several sub-block counts -> the assembled C must equal the
CPU oracle's product on EVERY rank (structure bit-exact, values in tolerance); row-sharded SpMV likewise.
Prints one line per rank and exits non-zero on any mismatch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sparse_matrix_b200 as S  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from sparse_matrix_b200 import distributed as D  # noqa: E402
from sparse_matrix_b200 import generators as G  # noqa: E402
from util import TOL, as_csr_matrix, random_csr  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = S.Handle(local)
    D.init_comm(h)
    info = h.comm_info()
    assert info["rank"] == rank and info["world"] == world
    rng = np.random.default_rng(3)          # same seed on every rank: identical inputs without a broadcast
    cases = [("rmat14", G.rmat(14, 12), np.float64), ("poisson128", G.poisson2d(128), np.float64),
             ("uniform_i64", G.uniform_random(5000, 7000, 9, seed=4, dtype=np.int64, int_range=99), np.int64),
             ("unsorted", random_csr(rng, 900, 900, rng.integers(0, 60, size=900), sorted_rows=False), np.float64)]
    checked = 0
    for name, m, dt in cases:
        b = m if m[0] == m[1] else G.transpose(m)
        off, idx, val = O.mul_hash(m, b, True)
        sabs = O.mul_hash(m[:4] + (np.abs(m[4]),), b[:4] + (np.abs(b[4]),), True)[2] if np.dtype(dt).kind == "f" else None
        dAfull = S.DeviceCsr.upload(as_csr_matrix(m, is_sorted=False), h)
        dB = dAfull if b is m else S.DeviceCsr.upload(as_csr_matrix(b, is_sorted=False), h)
        # rank 0's copy of the values replicated through the library (exercises spam_comm_broadcast)
        i = dAfull.info()
        h.comm_broadcast(i["d_val"], i["nnz"] * np.dtype(dt).itemsize, 0)
        starts, _ = dAfull.rows_to_parts(dB, world)
        r0, r1 = int(starts[rank]), int(starts[rank + 1])
        blk = dAfull.slice_rows(r0, r1)
        for nsub, mode in ((1, 0), (4, 0), (3, 1), (2, 2), (5, 2)):
            g = blk.matmul_gathered(dB, r0, m[0], nsub=nsub, mode=mode)
            c = g.download()
            g.free()
            assert np.array_equal(c.offsets, off), (name, nsub, mode, "row_ptr")
            assert np.array_equal(c.indices, idx), (name, nsub, mode, "col_idx")
            if sabs is None:
                assert np.array_equal(c.vals, val), (name, nsub, mode, "values")
            else:
                assert np.all(np.abs(c.vals - val) <= TOL[np.dtype(dt)] * sabs), (name, nsub, mode, "values")
            checked += 1
        if np.dtype(dt).kind == "f" and m[0] == m[1]:
            x = np.linspace(-1, 1, m[1])
            dx = torch.from_numpy(x).cuda()
            dy = torch.zeros(m[0], dtype=torch.float64, device="cuda")
            rows_of = [int(starts[r + 1] - starts[r]) for r in range(world)]
            h.synchronize(); torch.cuda.synchronize()
            blk.spmv_gathered(dx.data_ptr(), dy.data_ptr(), rows_of)
            h.synchronize()
            want = O.spmv(m[0], m[1], m[2], m[3], m[4], x)
            bound = 1e-12 * O.spmv(m[0], m[1], m[2], m[3], np.abs(m[4]), np.abs(x))
            assert np.all(np.abs(dy.cpu().numpy() - want) <= bound), (name, "spmv")
            checked += 1
        blk.free()
        if dB is not dAfull:
            dB.free()
        dAfull.free()
    # DOK -> CSR with the triplet stream spread over the ranks: rank r holds the r-th contiguous piece
    for dt in (np.int64, np.float64):
        u = G.uniform_random(20_000, 30_000, 8, seed=6, dtype=dt, int_range=1 << 12)
        tr, tc, tv = G.triplets_with_rewrites(u, seed=5, dup_frac=0.2, zero_frac=0.05)
        n = len(tv)
        lo, hi = n * rank // world, n * (rank + 1) // world
        d_r = torch.from_numpy(tr[lo:hi].view(np.int64).copy()).cuda()
        d_c = torch.from_numpy(tc[lo:hi].view(np.int64).copy()).cuda()
        d_v = torch.from_numpy(tv[lo:hi].copy()).cuda()
        torch.cuda.synchronize()
        blk, r0 = S.DeviceCsr.from_triplets_sharded(h, dt, u[0], u[1], hi - lo, d_r.data_ptr(), d_c.data_ptr(), d_v.data_ptr())
        got = blk.download()
        blk.free()
        off, idx, val = O.dok_to_csr(u[0], u[1], tr, tc, tv)
        per = -(-u[0] // world)
        assert r0 == min(u[0], per * rank), (r0, per, rank)
        r1 = min(u[0], r0 + per)
        e0, e1 = int(off[r0]), int(off[r1])
        assert got.rows() == r1 - r0
        assert np.array_equal(got.offsets, off[r0:r1 + 1] - off[r0]) and np.array_equal(got.indices, idx[e0:e1])
        assert np.array_equal(got.vals, val[e0:e1]), "sharded DOK -> CSR values"
        checked += 1
    print(f"rank {rank}/{world}: {checked} gathered results match the oracle; peer_mapped={h.comm_info()['peer_mapped']}", flush=True)
    h.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
