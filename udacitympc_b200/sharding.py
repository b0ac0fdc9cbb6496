"""Index sharding of one batch of independent MPC problems over devices / ranks: contiguous index
ranges, the remainder to the last one -- the rule b200mpc_solve_batch_multi (csrc/capi.cu) applies to its handles, and
the one bench.py applies to torchrun ranks.  No data-path collective: results are gathered by the host.

Also the rank-reduction of the timing contract (max over ranks of the device time), written against the
torch.distributed API so the same code runs on NCCL (bench.py) and on gloo (tests/test_sharding_gloo.py)."""


def shard_range(B, world, rank):
    """[lo, hi) of rank `rank` of `world`: B // world problems each, the last rank also takes the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need world >= 1 and 0 <= rank < world")
    per = B // world
    lo = per * rank
    hi = B if rank == world - 1 else per * (rank + 1)
    return lo, hi


def max_over_ranks(dist, t, world):
    """t: 1-element float64 tensor holding this rank's device time.  Returns (max over ranks, list of every rank's
    value); world == 1 needs no process group."""
    if world == 1:
        return float(t.item()), [float(t.item())]
    import torch
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    m = t.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(m.item()), [float(x.item()) for x in allt]


def throughput(total_problems_per_step, steps, max_ms):
    """whole-job solves/s: every rank's problems of every timed step over the slowest rank's device time"""
    return total_problems_per_step * steps / (max_ms * 1e-3)
