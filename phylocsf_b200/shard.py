"""Multi-GPU plumbing: alignments are independent, so the path shards by batch with no collective on
the data path. One process per GPU; each rank scores a contiguous, size-balanced slice
of the alignment list and rank 0 gathers the per-region results in input order (host side, after
the kernels). torch.distributed is used only for that gather and for the timing barrier."""
from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(weights: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous slices [lo, hi) of the item list, balanced by weight (e.g. codon columns per
    alignment). Contiguity keeps the output order a plain concatenation of the ranks' results."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        i = min(max(i, cuts[-1]), n)
        cuts.append(i)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_in_order(local: np.ndarray, rank: int, world: int, dst: int = 0):
    """Concatenate the ranks' result arrays (first axis) on rank `dst`, in rank order (= input order for
    contiguous shards). Returns the full array on dst, None elsewhere. Works on gloo and nccl groups."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    objs = [None] * world if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local), objs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate(objs, axis=0)
