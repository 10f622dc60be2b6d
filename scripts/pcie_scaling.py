"""
GPU box with N GPUs: what the host side can move when 1, 2, 4 ... N ranks copy at the same time - pinned H2D, D2H and both
at once, per rank and in aggregate, through this library's own plumbing calls (no torch on the data path).  It is the bus
ceiling the end-to-end number of bench.py (one h264gpu_compose_regions_host call per step: ~258 MB up, ~262 MB down per
rank) has to be read against.  Launch:  python scripts/pcie_scaling.py --gpus 8   -> one JSON object on stdout.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, n_active, barrier, q, mb):
    import __graft_entry__ as ge
    ge.load_package()
    import bench
    from h264_scroll_encoder_b200 import binding
    numa = bench.bind_to_gpu_numa_node(rank)
    gc = binding.GpuComposer(rank)
    L = gc.lib
    n = mb << 20
    h_in, h_out = L.h264gpu_host_alloc(n), L.h264gpu_host_alloc(n)
    ctypes.memset(h_in, 1, n); ctypes.memset(h_out, 2, n)
    d_in, d_out = gc.alloc(n), gc.alloc(n)
    s1, s2 = L.h264gpu_stream_create(), L.h264gpu_stream_create()
    res = {}
    for name, up, down in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
        for timed in (0, 1):
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(4):
                if up:
                    gc.check(L.h264gpu_memcpy_h2d(d_in.ptr, h_in, n, s1))
                if down:
                    gc.check(L.h264gpu_memcpy_d2h(h_out, d_out.ptr, n, s2))
            gc.check(L.h264gpu_stream_sync(s1)); gc.check(L.h264gpu_stream_sync(s2))
            dt = (time.perf_counter() - t0) / 4
        res[name] = n / dt / 1e9                      # GB/s per direction
    q.put((rank, res, numa))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mb", type=int, default=256)
    a = ap.parse_args()
    out = {"copy_mib": a.mb, "ranks": {}}
    counts = [c for c in (1, 2, 4, 8) if c <= a.gpus]
    ctx = mp.get_context("spawn")
    for c in counts:
        barrier, q = ctx.Barrier(c), ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, c, barrier, q, a.mb)) for r in range(c)]
        for p in ps:
            p.start()
        got = sorted(q.get() for _ in ps)
        for p in ps:
            p.join()
        e = {}
        for k in ("h2d", "d2h", "both"):
            v = [g[1][k] for g in got]
            e[k] = {"per_rank_min_gbs": min(v), "per_rank_mean_gbs": sum(v) / len(v), "aggregate_gbs": sum(v)}
        e["binding"] = [g[2] for g in got]
        out["ranks"][str(c)] = e
    print(json.dumps(out))


if __name__ == "__main__":
    main()
