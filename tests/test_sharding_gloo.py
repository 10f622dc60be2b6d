"""
N>1 host logic on CPU: two `gloo` ranks shard a batch of independent streams exactly as bench.py does on
two GPUs (block partition, no data-path collective), compose their shards with the CPU checker, and the
gathered result equals the single-process batch.  Also covers the max/sum reduction used for reporting.
"""
import hashlib
import os
import socket
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _digests(first, count, frames):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as ge
    ge.load_package()
    po = ge.load_oracle()
    import helpers
    from h264_scroll_encoder_b200 import workloads
    if count == 0:
        return [], 0
    b = workloads.make_config5(count, frames, stream0=first)
    hints = helpers.dense_hints(po, b)
    orc = po.Oracle()
    out = [hashlib.sha256(helpers.checker_frame(orc, b, hints, f).tobytes()).hexdigest() for f in range(b.n_frames)]
    return out, b.n_frames * b.n_mbs


def _worker(rank, world, port, n_streams, frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_package()
    from h264_scroll_encoder_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_streams(n_streams, rank, world)
    dig, mbs = _digests(first, count, frames)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, dig))
    times, units = sharding.reduce_job(dist, "cpu", [1.0 + rank, 5.0 - rank], [mbs])
    if rank == 0:
        q.put((gathered, times, units))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_streams_partition():
    from h264_scroll_encoder_b200 import sharding
    for n in (0, 1, 7, 128, 1024):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.shard_streams(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_streams(4, 2, 2)


def test_two_rank_gloo_shards_equal_single_process(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_streams, frames, world, port = 3, 2, 2, _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_streams, frames, q)) for r in range(world)]
    for p in ps:
        p.start()
    gathered, times, units = q.get(timeout=240)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole, mbs = _digests(0, n_streams, frames)
    joined = [d for _, dig in sorted(gathered) for d in dig]
    assert joined == whole                      # stream order == rank order, nothing exchanged
    assert times == [2.0, 5.0]                  # max over ranks
    assert units == [float(mbs)]                # sum over ranks
