"""-m "not gpu": the N>1 host logic bench.py runs under torchrun -- the library's index-range sharding rule
(udacitympc_b200/sharding.py, the same rule b200mpc_solve_batch_multi applies to its handles), the max-over-ranks
timing and the rank-0 aggregate -- with world_size 2 on the gloo backend.  The data path has no collective:
every rank "solves" its index range, the host gathers."""
import inspect
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from conftest import ROOT
from udacitympc_b200.sharding import max_over_ranks, shard_range, throughput


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, world, rank)
    mine = torch.arange(lo, hi, dtype=torch.float64) * 2.0        # stand-in for this rank's solves of problems [lo, hi)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)          # this rank's device time in ms
    tmax, per_rank = max_over_ranks(dist, t, world)
    # host gather of the shards (ragged: the last rank also holds the remainder)
    sizes = [shard_range(B, world, r)[1] - shard_range(B, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.float64)
    buf[:hi - lo] = mine
    gathered = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, buf)
    whole = torch.cat([g[:n] for g, n in zip(gathered, sizes)])
    if rank == 0:
        q.put((tmax, per_rank, lo, hi, bool((whole == torch.arange(B, dtype=torch.float64) * 2.0).all()), throughput(B, 20, tmax)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    B = 65537   # ragged on purpose
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, per_rank, lo, hi, whole_ok, value = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert tmax == 11.0 and per_rank == [10.0, 11.0] and (lo, hi) == (0, B // 2) and whole_ok
    assert value == B * 20 / 11e-3


def test_shard_ranges_cover_batch_exactly():
    for B in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(B, dtype=int)
            for r in range(world):
                lo, hi = shard_range(B, world, r)
                assert 0 <= lo <= hi <= B
                seen[lo:hi] += 1
            assert (seen == 1).all()


def test_bench_and_library_use_this_rule():
    """bench.py shards with the package's function; capi.cu's b200mpc_solve_batch_multi states the same rule."""
    import bench
    src = inspect.getsource(bench.run_ours)
    assert "from udacitympc_b200.sharding import" in src and "shard_range(B, world, rank)" in src
    c = open(os.path.join(ROOT, "udacitympc_b200", "csrc", "capi.cu")).read()
    assert "(long long)B / n_handles * g" in c and "g == n_handles - 1 ? B" in c
    assert bench.auto_streams(65536) == 6 and bench.auto_streams(8192) == 16 and bench.auto_streams(1 << 20) == 6
