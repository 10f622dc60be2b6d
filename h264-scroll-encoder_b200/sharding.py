"""
Multi-GPU partitioning of the path: streams are independent, so ranks take contiguous blocks
of streams and nothing is exchanged on the data path.  The only cross-rank traffic is the reduction of the
timing / unit counters for reporting (max of the times, sum of the units).
"""


def shard_streams(n_streams, rank, world):
    """Block partition: (first stream, count) of `rank`.  Earlier ranks take the remainder."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(n_streams, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_job(dist, device, step_seconds, units):
    """max over ranks of the per-step times, sum over ranks of the units (torch.distributed, any backend)."""
    import torch
    t = torch.tensor(list(step_seconds), dtype=torch.float64, device=device)
    u = torch.tensor(list(units), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in u]
