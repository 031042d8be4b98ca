"""Row-sharded SpGEMM across the GPUs of one node: one process per GPU, torch.distributed for the
plumbing (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).

The reference parallelises mul_hash over flop-balanced contiguous row blocks, one per thread
(spam_csr/src/mul_hash.rs:38-64), each writing a disjoint slice of the output (:121-128).  Across
GPUs the same split is used with tnum = world size: rank r multiplies rows
[row_starts[r], row_starts[r+1]) of A by a replicated B.  Rows of C depend only on the matching rows of
A, so the product itself needs no data-path collective; the collectives are
  * broadcast of B (and A) from rank 0                       -- replicate()
  * all-gather of the per-rank nnz(C) (world x u64)          -- to offset-fix each shard's row_ptr
  * all-gather-v of row_ptr / col_idx / val shards           -- NCCL has no AllGatherv:
                                                                one broadcast per rank, batched.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def init_comm(handle) -> None:
    """Collective.  Gives `handle` its in-library communicator (spam_comm_init): rank 0 makes the 128-byte id,
    torch.distributed (any backend) carries it to the other ranks — the one thing the library cannot do itself."""
    from . import csr
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [csr.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    handle.comm_init(uid[0], rank, world)


def partition_rows_from_flops(flop_per_row: np.ndarray, parts: int) -> np.ndarray:
    """rows_to_threads partition (mul_hash.rs:51-62) on the host: used by the CPU tests to check the
    device routine spam_rows_to_parts, and to split work when the flop vector is already on the host."""
    ps = np.concatenate([[0], np.cumsum(flop_per_row.astype(np.uint64))]).astype(np.uint64)
    total = int(ps[-1])
    avg = -(-total // parts)
    starts = [0]
    for t in range(1, parts):
        starts.append(int(np.searchsorted(ps, np.uint64(avg * t), side="right")) - 1)
    starts.append(len(flop_per_row))
    return np.asarray(starts, dtype=np.uint64)


def row_cost(flop_per_row: np.ndarray) -> np.ndarray:
    """Host mirror of the device-time estimate behind spam_rows_to_parts_cost (api.cu row_cost_q):
    products x a per-product weight (1/16 units) that depends on the size class of the row."""
    f = flop_per_row.astype(np.uint64)
    w = np.select([f <= 128, f <= 8192], [18, 16], 64)
    return np.minimum(f * w.astype(np.uint64), np.uint64(0xFFFFFFFF))


def replicate(tensors: Sequence[torch.Tensor], src: int = 0) -> float:
    """Broadcast the CSR arrays of B (and A) from `src` to every rank.  Returns seconds (host clock
    around a device sync; setup cost, reported separately from the product)."""
    import time
    if tensors and tensors[0].is_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in tensors:
        dist.broadcast(t, src=src)
    if tensors and tensors[0].is_cuda:
        torch.cuda.synchronize()
    return time.perf_counter() - t0


def gather_counts(local_count: int, device) -> List[int]:
    """all-gather of one u64 per rank (nnz of each rank's C shard)."""
    world = dist.get_world_size()
    mine = torch.tensor([local_count], dtype=torch.int64, device=device)
    out = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    return [int(x.item()) for x in out]


def all_gather_v(shard: torch.Tensor, counts: Sequence[int], out: torch.Tensor = None) -> torch.Tensor:
    """all-gather-v: rank r contributes `counts[r]` elements; every rank ends with the concatenation.
    One broadcast per source rank into its slice of the output (batched in one NCCL group on CUDA)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    total = int(sum(counts))
    if out is None:
        out = torch.empty(total, dtype=shard.dtype, device=shard.device)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    out[offs[rank]:offs[rank + 1]].copy_(shard[:counts[rank]])
    works = []
    for r in range(world):
        if counts[r] == 0:
            continue
        works.append(dist.broadcast(out[offs[r]:offs[r + 1]], src=r, async_op=True))
    for w in works:
        w.wait()
    return out


def offset_fixed_row_ptr(local_ptr: torch.Tensor, nnz_before: int) -> torch.Tensor:
    """A rank's row_ptr shard (rows+1 entries starting at 0) -> global offsets, without its leading 0
    except on rank 0, so that the shards concatenate to the full row_ptr."""
    fixed = local_ptr + nnz_before
    return fixed if dist.get_rank() == 0 else fixed[1:]


def gathered_csr(local_ptr: torch.Tensor, local_idx: torch.Tensor, local_val: torch.Tensor,
                 rows_per_rank: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, List[int]]:
    """Assemble the full C on every rank from the row shards."""
    counts = gather_counts(int(local_idx.shape[0]), local_idx.device)
    rank = dist.get_rank()
    before = int(sum(counts[:rank]))
    ptr_shard = offset_fixed_row_ptr(local_ptr, before)
    ptr_counts = [rows_per_rank[r] + (1 if r == 0 else 0) for r in range(len(rows_per_rank))]
    ptr = all_gather_v(ptr_shard.contiguous(), ptr_counts)
    idx = all_gather_v(local_idx, counts)
    val = all_gather_v(local_val, counts)
    return ptr, idx, val, counts
