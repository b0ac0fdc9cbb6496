"""-m "not gpu": the N>1 host logic of bench.py (index-range sharding, max-over-ranks timing, rank-0 reporting)
with world_size 2 on the gloo backend.  The data path has no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp


def shard_range(B, world, rank):
    """contiguous index ranges, remainder to the last rank -- the rule b200mpc_solve_batch_multi uses"""
    lo = B // world * rank
    hi = B if rank == world - 1 else B // world * (rank + 1)
    return lo, hi


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, world, rank)
    mine = torch.arange(lo, hi, dtype=torch.float64) * 2.0        # stand-in for this rank's solves
    t = torch.tensor([10.0 + rank])                               # this rank's device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(hi - lo)])
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    gathered = [torch.zeros(shard_range(B, world, r)[1] - shard_range(B, world, r)[0], dtype=torch.float64) for r in range(world)]
    dist.all_gather(gathered, mine) if all(g.numel() == mine.numel() for g in gathered) else None
    if rank == 0:
        q.put((float(t.item()), float(n.item()), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    B = 65536
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, total, lo, hi = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert tmax == 11.0 and total == B and (lo, hi) == (0, B // 2)


def test_shard_ranges_cover_batch_exactly():
    for B in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(B, dtype=int)
            for r in range(world):
                lo, hi = shard_range(B, world, r)
                assert 0 <= lo <= hi <= B
                seen[lo:hi] += 1
            assert (seen == 1).all()
