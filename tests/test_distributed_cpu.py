"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU tensors: the flop-balanced row
partition, the offset-fixed row_ptr shards and the all-gather-v assembly of C.  The per-rank product is
played by the oracle here (no GPU in this container); on the GPU box the same routines run over NCCL
(bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import random_csr


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as O
        from sparse_matrix_b200 import distributed as D
        rng = np.random.default_rng(seed)
        rows = 257
        a = random_csr(rng, rows, rows, rng.integers(0, 30, size=rows) ** 2 % 37)   # skewed row lengths
        # rank 0 owns the input; everyone receives it (replicate = broadcast of the CSR arrays)
        if rank == 0:
            t_ptr = torch.from_numpy(a[2].view(np.int64).copy())
            t_idx = torch.from_numpy(a[3].view(np.int64).copy())
            t_val = torch.from_numpy(a[4].copy())
        else:
            t_ptr = torch.zeros(rows + 1, dtype=torch.int64)
            t_idx = torch.zeros(len(a[3]), dtype=torch.int64)
            t_val = torch.zeros(len(a[3]), dtype=torch.float64)
        D.replicate([t_ptr, t_idx, t_val], src=0)
        ptr, idx, val = t_ptr.numpy().view(np.uint64), t_idx.numpy().view(np.uint64), t_val.numpy()
        assert np.array_equal(ptr, a[2]) and np.array_equal(idx, a[3]) and np.array_equal(val, a[4])
        # flop-balanced partition == rows_to_threads with tnum = world (mul_hash.rs:51-62)
        flop, ro = O.rows_to_threads(rows, ptr, idx, ptr, world)
        starts = D.partition_rows_from_flops(flop, world)
        assert np.array_equal(starts, ro), (starts, ro)
        r0, r1 = int(starts[rank]), int(starts[rank + 1])
        # this rank's row block of A times the replicated B (the oracle stands in for the GPU product)
        blk = (r1 - r0, rows, ptr[r0:r1 + 1] - ptr[r0], idx[int(ptr[r0]):int(ptr[r1])], val[int(ptr[r0]):int(ptr[r1])])
        if r1 > r0:
            c_off, c_idx, c_val = O.mul_hash(blk, (rows, rows, ptr, idx, val), True)
        else:
            c_off, c_idx, c_val = np.zeros(1, np.uint64), np.zeros(0, np.uint64), np.zeros(0)
        rows_per = [int(starts[r + 1] - starts[r]) for r in range(world)]
        g_ptr, g_idx, g_val, counts = D.gathered_csr(torch.from_numpy(c_off.view(np.int64).copy()),
                                                     torch.from_numpy(c_idx.view(np.int64).copy()),
                                                     torch.from_numpy(c_val.copy()), rows_per)
        full = O.mul_hash((rows, rows, ptr, idx, val), (rows, rows, ptr, idx, val), True)
        assert sum(counts) == len(full[1])
        assert np.array_equal(g_ptr.numpy().view(np.uint64), full[0])
        assert np.array_equal(g_idx.numpy().view(np.uint64), full[1])
        assert np.array_equal(g_val.numpy(), full[2])        # same per-row order everywhere: bit-identical
        # all_gather_v with an empty contribution
        mine = torch.arange(rank * 10, rank * 10 + (0 if rank == 1 else 3 + rank), dtype=torch.int64)
        cnts = [3, 0, 5][:world]
        got = D.all_gather_v(mine, cnts)
        want = np.concatenate([np.arange(r * 10, r * 10 + cnts[r]) for r in range(world)])
        assert np.array_equal(got.numpy(), want)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_product_over_gloo(world, tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 1234 + world, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").exists()


def test_partition_edge_cases(oracle):
    from sparse_matrix_b200 import distributed as D
    for flops, parts in [(np.zeros(5, np.uint64), 3), (np.array([7], np.uint64), 4), (np.array([0, 0, 9, 0], np.uint64), 2),
                         (np.arange(100, dtype=np.uint64), 8), (np.array([10**6, 1, 1, 1], np.uint64), 4)]:
        starts = D.partition_rows_from_flops(flops, parts)
        assert starts[0] == 0 and starts[-1] == len(flops) and np.all(np.diff(starts.astype(np.int64)) >= 0)
        # the oracle's rows_to_threads on a matrix with exactly these per-row flops: A = diag-like
        # selector of B rows whose lengths are `flops`
        n = len(flops)
        a_off = np.arange(n + 1, dtype=np.uint64)
        a_idx = np.arange(n, dtype=np.uint64)
        b_off = np.concatenate([[0], np.cumsum(flops)]).astype(np.uint64)
        f, ro = oracle.rows_to_threads(n, a_off, a_idx, b_off, parts)
        assert np.array_equal(f, flops) and np.array_equal(ro, starts)


def test_cost_balanced_partition_host_mirror():
    """spam_rows_to_parts_cost's host mirror: cost = products x a weight between 16 and 64 sixteenths, uniform
    rows give the flop partition, and on a power-law flop vector the heaviest block holds fewer products."""
    from sparse_matrix_b200 import distributed as D
    f = np.arange(0, 40_000, 7, dtype=np.uint64)
    c = D.row_cost(f)
    assert c[0] == 0 and np.all(c >= 16 * f) and np.all(c <= 64 * f)
    assert D.row_cost(np.array([2**31], np.uint64))[0] == 0xFFFFFFFF       # saturates like the device u32
    uni = np.full(1000, 25, np.uint64)
    assert np.array_equal(D.partition_rows_from_flops(uni, 8), D.partition_rows_from_flops(D.row_cost(uni), 8))
    rng = np.random.default_rng(3)
    pl = np.sort((rng.pareto(1.2, 50_000) * 60).astype(np.uint64) + 1)[::-1].copy()   # heavy rows first, like R-MAT
    by_f = D.partition_rows_from_flops(pl, 8).astype(np.int64)
    by_c = D.partition_rows_from_flops(D.row_cost(pl), 8).astype(np.int64)
    assert by_c[1] < by_f[1]                         # the first (heaviest) block shrinks
    cost = D.row_cost(pl).astype(np.float64)
    blocks = lambda st: np.array([cost[st[i]:st[i + 1]].sum() for i in range(8)])
    assert blocks(by_c).max() <= blocks(by_f).max()
