"""N>1 host logic on CPU: world_size-2 gloo processes shard an alignment list, 'score' their slice and
rank 0 gathers the results in input order (no collective on the data path)."""
import os
import socket

import numpy as np
import pytest

from phylocsf_b200 import shard


def test_shard_bounds_balanced_and_contiguous():
    w = [100, 99, 99] * 50 + [5000]
    for world in (1, 2, 3, 4, 8):
        b = shard.shard_bounds(w, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(w)
        for (l0, h0), (l1, h1) in zip(b, b[1:]):
            assert h0 == l1 and l0 <= h0
        loads = [sum(w[lo:hi]) for lo, hi in b]
        assert max(loads) <= sum(w) / world + max(w)
    assert shard.shard_bounds([], 4) == [(0, 0)] * 4
    assert shard.shard_bounds([1, 1], 4)[-1][1] == 2


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = [10 + (i % 7) for i in range(101)]
    lo, hi = shard.shard_bounds(weights, world)[rank]
    local = np.array([[i, weights[i] * 2.5] for i in range(lo, hi)], dtype=np.float64)  # stand-in for (region, score)
    dist.barrier()
    full = shard.gather_in_order(local, rank, world)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_preserves_input_order():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.shape == (101, 2)
    assert (full[:, 0] == np.arange(101)).all()
    assert (full[:, 1] == np.array([(10 + (i % 7)) * 2.5 for i in range(101)])).all()
