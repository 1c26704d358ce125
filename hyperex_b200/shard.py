"""Record sharding for the one-process-per-GPU layout.

Records never interact (the loop at utils.rs:270 carries no state), so a batch is cut into contiguous
input-order ranges balanced by bases, one per rank; every rank scans its own range on its own GPU and
rank 0 concatenates the per-rank result tables by rank index, which restores input order.  No collective
sits on the data path; the final gather of the (small) result tables is host-side plumbing done with
torch.distributed (gloo on CPU tensors, so it also runs without GPUs in the tests).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def shard_ranges(offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """[(r0, r1)] per rank: contiguous, input order, ~equal bases; ranks may be empty when n < world"""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    base, total = int(offsets[0]), int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = base + total * r // world
        # first record whose end passes the target
        c = int(np.searchsorted(offsets[1:], np.uint64(target), side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def shard_view(arena: np.ndarray, offsets: np.ndarray, r0: int, r1: int):
    """(arena slice, offsets slice) of records [r0, r1) without copying; offsets stay absolute-relative
    to the slice start"""
    o = np.asarray(offsets, dtype=np.uint64)
    lo, hi = int(o[r0]), int(o[r1])
    return arena[lo:hi], (o[r0:r1 + 1] - np.uint64(lo)).astype(np.uint64)


def run_sharded(compute: Callable[[np.ndarray, np.ndarray], np.ndarray], arena: np.ndarray, offsets: np.ndarray,
                rank: int, world: int, n_pairs: int, dtype, group=None):
    """Every rank runs `compute` (normally Context.run_batch) on its range; rank 0 returns the full
    (n_records, n_pairs) table in input order, other ranks return None."""
    import torch
    import torch.distributed as dist

    ranges = shard_ranges(offsets, world)
    r0, r1 = ranges[rank]
    a, o = shard_view(arena, offsets, r0, r1)
    local = compute(a, o) if r1 > r0 else np.zeros((0, n_pairs), dtype=dtype)
    local = np.ascontiguousarray(local).reshape(r1 - r0, n_pairs)
    if world == 1:
        return local
    itemsize = np.dtype(dtype).itemsize
    sizes = [(b - a_) * n_pairs * itemsize for a_, b in ranges]
    mine = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    if rank == 0:
        bufs = [torch.empty(s, dtype=torch.uint8) for s in sizes]
        bufs[0].copy_(mine)
        for src in range(1, world):
            if sizes[src]:
                dist.recv(bufs[src], src=src, group=group)
        flat = torch.cat(bufs).numpy()
        return flat.view(dtype).reshape(len(offsets) - 1, n_pairs)
    if sizes[rank]:
        dist.send(mine, dst=0, group=group)
    return None
