#!/usr/bin/env python
"""
bench.py - composed macroblocks/second of bit-exact P-frames (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a, through the C ABI)
  python bench.py --impl reference ...                     the reference's own CPU code on the host cores

Workload (config.workload): BASELINE.json configs[4] "batched 1024 independent 1080p UI streams sharded
across 8xB200": every GPU takes 128 streams x 64 frames of the config-3 layout (static chrome, three
carousels, a two-reference vertical grid, a 25x25-MB spliced dynamic region), so the
8-GPU run is exactly the 1024-stream batch and per-GPU work is fixed (weak scaling, no collective).

A step = one pass of the hot path over the rank's whole batch.
  value   MB/s with all inputs resident in HBM (dense hint maps, payload, descriptors): the compose kernel
  e2e     MB/s through the host-buffer path: pinned host region lists + payload -> H2D -> rasterise ->
          compose -> pack -> D2H of the packed Annex-B stream, every step
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "composed_macroblocks_per_second_bit_exact_p_frames"
UNIT = "MB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~0.5 ms during the timed region
    (the region lasts tens of milliseconds: too short for nvidia-smi's 100 ms loop)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.stop_flag, self.t, self.max = index, [], 0, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                       # noqa: BLE001
            self.nv, self.err = None, str(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:                        # noqa: BLE001
                pass
            time.sleep(0.0005)

    def start(self):
        if self.nv:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.t.join(timeout=1)
        nv, names = self.nv, []
        for name, bit in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                          ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                          ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                          ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            b = getattr(nv, bit, None)
            if b is None:
                b = getattr(nv, bit.replace("ClocksEventReason", "ClocksThrottleReason"), 0)
            if self.reasons & int(b):
                names.append(name)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "samples": len(self.sm), "reasons": names}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs of the NUMA node GPU `index` hangs off, BEFORE any pinned memory is allocated:
    the end-to-end path moves ~0.5 GB per step per GPU through pinned host buffers, and with several ranks on a
    two-socket host buffers on the wrong node cross the socket link twice.  Returns the node (or None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                                             # noqa: BLE001
        return None


# --------------------------------------------------------------------------------------- reference arm

def cpu_sample(po, streams=2, frames=8):
    """Bounded sample of the same workload as dense hints for the CPU code."""
    from h264_scroll_encoder_b200 import workloads, abi
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    b = workloads.make_config5(streams, frames)
    hints = np.zeros((b.n_frames, b.n_mbs), abi.HINT)
    for f in range(b.n_frames):
        r0, rc = int(b.rdesc["region_first"][f]), int(b.rdesc["region_count"][f])
        hints[f] = po.raster_regions_numpy(b.regions[r0:r0 + rc], b.mb_w, b.mb_h, b.dyn_info[int(b.rdesc["dyn_info_off"][f]):])
    # per-frame tables re-based on one shared payload buffer
    tab = np.zeros((b.n_frames, 626), np.uint32)
    for f in range(b.n_frames):
        t0 = int(b.dyn["table_off"][f])
        tab[f] = b.dyn_tab[t0:t0 + 626] + np.uint32(8 * int(b.dyn["payload_off"][f]))
    return b, np.ascontiguousarray(hints), np.ascontiguousarray(tab), b.payload


def run_cpu(po, threads, target_seconds, kind=None):
    """Time the reference's CPU path (oracle/_ref when present, else the oracle port) on `threads` threads."""
    use_ref = po.have_ref() if kind is None else (kind == "reference")
    chk = po.Ref() if use_ref else po.Oracle()
    b, hints, tab, payload = cpu_sample(po)
    hdr = b.desc["hdr_bits"][0].copy()
    nb = int(b.desc["hdr_nbits"][0])
    # calibrate reps on one thread
    t1, _ = chk.time_hinted(hints, b.n_frames, b.mb_w, b.mb_h, 2, 0, hdr, nb, payload, tab, 626, 1)
    reps = max(1, int(target_seconds / max(t1, 1e-6)))
    res = [None] * threads

    def work(i):
        res[i] = chk.time_hinted(hints, b.n_frames, b.mb_w, b.mb_h, 2, 0, hdr, nb, payload, tab, 626, reps)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    mbs = threads * reps * b.n_frames * b.n_mbs
    return {"value": mbs / wall, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": "%d threads x %d passes over %d frames (2 streams x 8 frames of the config-5 workload, "
                      "%d MBs/frame, dense hints), %.1f s wall" % (threads, reps, b.n_frames, b.n_mbs, wall)}, wall, reps


def main_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    po = ge.load_oracle()
    ge.load_package()
    threads = os.cpu_count() or 1
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        r, wall, reps = run_cpu(po, threads, target_seconds=max(1.0, 20.0 / max(args.steps + args.warmup, 1)))
        if i >= args.warmup:
            vals.append(r["value"]); walls.append(wall)
        last = r
    v = float(np.mean(vals))
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args), "cpu_baseline": last,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def workload_config(args):
    return {"workload": "config5: batched independent 1080p UI streams, %d streams x %d frames per GPU "
                        "(1920x1088 = 120x68 MBs, config-3 layout with a 25x25-MB spliced dynamic region, "
                        "num_refs 2, skip OFF, reference predictor)" % (args.streams, args.frames),
            "streams_per_gpu": args.streams, "frames_per_stream": args.frames, "mbs_per_frame": 8160,
            "l2_policy": "inputs larger than L2: %.0f MB of hints + payload per GPU per step vs 126 MB L2"
                         % ((args.streams * args.frames * (8160 * 8 + 23500)) / 1e6),
            "sharding": "streams block-partitioned over ranks, no collective"}


# --------------------------------------------------------------------------------------------- our arm

def main_ours(args):
    rank, local, world = dist_env()
    ge.load_package()
    from h264_scroll_encoder_b200 import binding, workloads, abi
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa_node = None if args.no_numa_bind else bind_to_gpu_numa_node(local)
    gc = binding.GpuComposer(local)      # raises if the CUDA library or a B200 is missing: no fallback
    L = gc.lib

    b = workloads.make_config5(args.streams, args.frames, stream0=rank * args.streams)
    n, n_mbs = b.n_frames, b.n_mbs
    total_mbs = n * n_mbs

    # ---- device-resident inputs -------------------------------------------------------------------
    pay_pad = np.concatenate([b.payload, np.zeros(16, np.uint8)])
    d_desc, d_rdesc, d_reg = gc.to_device(b.desc), gc.to_device(b.rdesc), gc.to_device(b.regions)
    d_info, d_dyn, d_tab, d_pay = gc.to_device(b.dyn_info), gc.to_device(b.dyn), gc.to_device(b.dyn_tab), gc.to_device(pay_pad)
    d_hints = gc.alloc(8 * b.hint_stride * n)
    d_out = gc.alloc(b.slot_cap * n)
    d_sizes = gc.alloc(4 * n)
    gc.raster_regions(d_rdesc, n, n_mbs, d_reg, d_info, d_hints)
    gc.sync()

    ev0, ev1 = L.h264gpu_event_create(), L.h264gpu_event_create()

    def barrier():
        gc.sync()
        if dist is not None:
            dist.barrier()
        gc.sync()

    def kernel_step():
        gc.compose_batch(d_desc, n, d_hints, d_dyn, d_tab, d_pay, d_out, d_sizes)

    for _ in range(args.warmup):
        kernel_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = gc.launch_count()
    gc.check(L.h264gpu_event_record(ev0, None))
    for _ in range(args.steps):
        kernel_step()
    gc.check(L.h264gpu_event_record(ev1, None))
    ms = L.h264gpu_event_elapsed_ms(ev0, ev1)
    launches = gc.launch_count() - launches0
    samples_timed = len(sampler.sm)
    # the timed region is ~10 ms - one or two NVML reads; the same kernel keeps running (untimed) while the sampler goes on,
    # so that the median is taken over a few dozen samples under the very same load
    t_end = time.perf_counter() + 0.06
    while time.perf_counter() < t_end:
        kernel_step()
        gc.sync()
    barrier()
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = samples_timed

    # ---- the rasteriser on its own: region lists -> dense hint maps, output-bound (8 B/MB written) ----
    for _ in range(2):
        gc.raster_regions(d_rdesc, n, n_mbs, d_reg, d_info, d_hints)
    gc.check(L.h264gpu_event_record(ev0, None))
    for _ in range(5):
        gc.raster_regions(d_rdesc, n, n_mbs, d_reg, d_info, d_hints)
    gc.check(L.h264gpu_event_record(ev1, None))
    raster_ms = L.h264gpu_event_elapsed_ms(ev0, ev1) / 5
    raster_bytes = 8 * total_mbs + b.regions.nbytes + b.rdesc.nbytes + b.dyn_info.nbytes     # map written, lists and dyn-info read

    if hasattr(L, "h264gpu_debug_stage_cycles"):                  # timing-instrumented experiment build only
        cyc = (ctypes.c_uint64 * 16)()
        L.h264gpu_debug_stage_cycles(cyc)
        tot = float(sum(cyc[:8])) or 1.0
        sys.stderr.write("stage cycles (thread 0 of every CTA, warm-up + timed): " +
                         " ".join("%d:%.1f%%" % (i, 100.0 * c / tot) for i, c in enumerate(cyc)) + "\n")
    sizes = d_sizes.download(np.uint32)
    ablate = bool(os.environ.get("H264SCROLL_BENCH_NOCHECK"))      # timing experiments on deliberately incomplete kernels
    assert ablate or ((sizes <= b.slot_cap).all() and (sizes > 5).all()), "a frame overflowed its slot"
    s_out = int(sizes.sum())
    s_dyn = int(sum(int(b.dyn_tab[int(b.dyn["table_off"][f]) + 625]) for f in range(n)) // 8)
    b_alg = 8 * total_mbs + s_dyn + 4 * 626 * n + s_out + 64 * n

    # ---- end to end: host buffers in, packed Annex-B stream out, every step ------------------------
    # The batch goes through in chunks of frames on a few streams, so that the H2D of one chunk, the kernels of
    # another and the D2H of a third overlap (PCIe is full duplex).  Every byte of input still crosses the bus
    # every step and every byte of the packed stream comes back.
    def pinned(a):
        a = np.ascontiguousarray(a)
        p = L.h264gpu_host_alloc(max(a.nbytes, 16))
        v = np.ctypeslib.as_array((ctypes.c_uint8 * max(a.nbytes, 1)).from_address(p))
        v[:a.nbytes] = a.view(np.uint8).reshape(-1)
        return p
    n_chunks = max(1, min(args.e2e_chunks, n))
    bounds = [n * c // n_chunks for c in range(n_chunks + 1)]

    def edge(arr_off, item, f, total):        # byte offset where frame f's share of a per-frame-indexed array starts
        return int(arr_off[f]) * item if f < n else total
    arrays = [  # (host array, device buffer, byte range of frames [f0, f1))
        (b.desc, d_desc, lambda f: f * b.desc.itemsize),
        (b.rdesc, d_rdesc, lambda f: f * b.rdesc.itemsize),
        (b.regions, d_reg, lambda f: edge(b.rdesc["region_first"], b.regions.itemsize, f, b.regions.nbytes)),
        (b.dyn_info, d_info, lambda f: edge(b.rdesc["dyn_info_off"], b.dyn_info.itemsize, f, b.dyn_info.nbytes)),
        (b.dyn, d_dyn, lambda f: f * b.dyn.itemsize),
        (b.dyn_tab, d_tab, lambda f: edge(b.dyn["table_off"], 4, f, b.dyn_tab.nbytes)),
        (b.payload, d_pay, lambda f: edge(b.dyn["payload_off"], 1, f, b.payload.nbytes)),
    ]
    ups = [(pinned(a), d, rng) for a, d, rng in arrays]
    h2d_bytes = sum(a.nbytes for a, _, _ in arrays)
    chunk_frames = max(bounds[c + 1] - bounds[c] for c in range(n_chunks))
    chunk_cap = b.slot_cap * chunk_frames // 2
    d_packed_c = gc.alloc(chunk_cap * n_chunks)
    d_offs_c = gc.alloc(8 * (chunk_frames + 1) * n_chunks)
    h_out = L.h264gpu_host_alloc(b.slot_cap * n // 2)
    h_offs = L.h264gpu_host_alloc(8 * (chunk_frames + 1) * n_chunks)
    offs_view = np.ctypeslib.as_array((ctypes.c_uint64 * ((chunk_frames + 1) * n_chunks)).from_address(h_offs))
    streams = [L.h264gpu_stream_create() for _ in range(min(args.e2e_streams, n_chunks))]
    LAG = args.e2e_lag

    zero_copy = args.e2e_zero_copy      # measured slower than the copy engine (8.8 vs 7.3 ms per step): off by default
    h_out_c = L.h264gpu_host_alloc(chunk_cap * n_chunks) if zero_copy else None

    def e2e_step_zero_copy():
        # the pack kernel writes the Annex-B stream of each chunk (and its frame offsets) straight into pinned host
        # memory: the device->host transfer happens inside the kernel, no size has to come back before it starts
        for c in range(n_chunks):
            f0, f1, st = bounds[c], bounds[c + 1], streams[c % len(streams)]
            for p, d, rng in ups:
                o0, o1 = rng(f0), rng(f1)
                if o1 > o0:
                    gc.check(L.h264gpu_memcpy_h2d(d.ptr + o0, p + o0, o1 - o0, st))
            gc.check(L.h264gpu_raster_regions(d_rdesc.ptr + f0 * b.rdesc.itemsize, f1 - f0, n_mbs, d_reg.ptr, d_info.ptr,
                                              d_hints.ptr, st))
            gc.check(L.h264gpu_compose_batch(d_desc.ptr + f0 * b.desc.itemsize, f1 - f0, d_hints.ptr, d_dyn.ptr, d_tab.ptr,
                                             d_pay.ptr, d_out.ptr, d_sizes.ptr + 4 * f0, st))
            oo = 8 * (chunk_frames + 1) * c
            gc.check(L.h264gpu_pack_frames(d_desc.ptr + f0 * b.desc.itemsize, d_sizes.ptr + 4 * f0, f1 - f0, d_out.ptr,
                                           h_out_c + chunk_cap * c, h_offs + oo, st))
        for st in streams:
            gc.check(L.h264gpu_stream_sync(st))
        total = 0
        for c in range(n_chunks):
            tot = int(offs_view[(chunk_frames + 1) * c + (bounds[c + 1] - bounds[c])])
            assert tot <= chunk_cap, "packed chunk larger than its buffer"
            total += tot
        return total

    def e2e_step_staged():
        total = 0
        for c in range(n_chunks + LAG):
            if c < n_chunks:
                f0, f1, st = bounds[c], bounds[c + 1], streams[c % len(streams)]
                for p, d, rng in ups:
                    o0, o1 = rng(f0), rng(f1)
                    if o1 > o0:
                        gc.check(L.h264gpu_memcpy_h2d(d.ptr + o0, p + o0, o1 - o0, st))
                gc.check(L.h264gpu_raster_regions(d_rdesc.ptr + f0 * b.rdesc.itemsize, f1 - f0, n_mbs, d_reg.ptr, d_info.ptr,
                                                  d_hints.ptr, st))
                gc.check(L.h264gpu_compose_batch(d_desc.ptr + f0 * b.desc.itemsize, f1 - f0, d_hints.ptr, d_dyn.ptr, d_tab.ptr,
                                                 d_pay.ptr, d_out.ptr, d_sizes.ptr + 4 * f0, st))
                oo = 8 * (chunk_frames + 1) * c
                gc.check(L.h264gpu_pack_frames(d_desc.ptr + f0 * b.desc.itemsize, d_sizes.ptr + 4 * f0, f1 - f0, d_out.ptr,
                                               d_packed_c.ptr + chunk_cap * c, d_offs_c.ptr + oo, st))
                gc.check(L.h264gpu_memcpy_d2h(h_offs + oo, d_offs_c.ptr + oo, 8 * (f1 - f0 + 1), st))
            d = c - LAG
            if d >= 0:
                st = streams[d % len(streams)]
                gc.check(L.h264gpu_stream_sync(st))
                tot = int(offs_view[(chunk_frames + 1) * d + (bounds[d + 1] - bounds[d])])
                assert tot <= chunk_cap, "packed chunk larger than its buffer"
                gc.check(L.h264gpu_memcpy_d2h(h_out + total, d_packed_c.ptr + chunk_cap * d, tot, st))
                total += tot
        for st in streams:
            gc.check(L.h264gpu_stream_sync(st))
        return total

    # The same pipeline behind ONE C-ABI call (h264gpu_compose_regions_host): the library cuts the batch into chunks and
    # drives streams, events and copies itself, so no Python sits between the enqueues.  This is the call a host
    # application makes; the two functions above are kept for comparison (--e2e-python).
    hb, hb_keep = gc.make_host_batch(b, pinned=True)
    h_stream_cap = b.slot_cap * n // 2
    h_offs_all = L.h264gpu_host_alloc(8 * (n + 1))
    h_sizes_all = L.h264gpu_host_alloc(4 * n)
    offs_all_view = np.ctypeslib.as_array((ctypes.c_uint64 * (n + 1)).from_address(h_offs_all))

    def e2e_step_capi():
        gc.compose_regions_host(hb, h_out, h_stream_cap, h_offs_all, h_sizes_all, n_chunks)
        return int(offs_all_view[n])

    e2e_step = e2e_step_zero_copy if zero_copy else (e2e_step_staged if args.e2e_python else e2e_step_capi)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        d2h = e2e_step()
    gc.sync()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert ablate or d2h == s_out, "packed stream size differs from the sum of the frame sizes"
    e2e_launches = 5 * n_chunks                                  # raster + compose (2 instantiations) + pack (2 kernels) per chunk
    if e2e_step is e2e_step_capi and not ablate:
        # the stream the call returned is the concatenation of the slots the kernel-only run left on the device
        fo = np.array(offs_all_view, dtype=np.uint64)
        assert np.array_equal(fo[1:] - fo[:-1], sizes.astype(np.uint64)), "end-to-end frame sizes differ from the kernel-only run"
        hv = np.ctypeslib.as_array((ctypes.c_uint8 * d2h).from_address(h_out))
        for f in (0, n // 2, n - 1):
            slot = d_out.download(np.uint8, count=int(sizes[f]), offset=int(b.desc["out_off"][f]))
            assert np.array_equal(hv[int(fo[f]):int(fo[f + 1])], slot), "end-to-end stream differs from the device slots"

    # ---- reduce over ranks (max time) ------------------------------------------------------------
    ms_step, e2e_step_s = ms / args.steps, e2e_s
    if dist is not None:
        import torch
        t = torch.tensor([ms_step, e2e_step_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step_s = float(t[0]), float(t[1])
        tm = torch.tensor([float(total_mbs), float(b_alg)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.SUM)
        job_mbs = float(tm[0])
    else:
        job_mbs = float(total_mbs)

    if rank == 0:
        peak, peak_kind = peaks()
        achieved = b_alg / (ms / args.steps * 1e-3) / 1e9      # this rank's kernel, GB/s
        line = {
            "metric": METRIC, "value": job_mbs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": job_mbs / e2e_step_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h + 8 * (n + n_chunks) + (4 * n if e2e_step is e2e_step_capi else 0)), "ms_per_step": 1e3 * e2e_step_s,
                    "chunks": n_chunks, "gpu_launches_per_step": e2e_launches, "host_numa_node": numa_node,
                    "path": ("pinned host region lists + payload -> H2D -> raster -> compose -> pack writes the Annex-B stream "
                             "into pinned host memory (D2H inside the kernel)" if zero_copy else
                             ("one h264gpu_compose_regions_host call: " if e2e_step is e2e_step_capi else "") +
                             "pinned host region lists + payload -> H2D -> raster -> compose -> pack -> D2H packed stream")
                            + ", in %d frame chunks over %d streams" % (n_chunks, 3 if e2e_step is e2e_step_capi else len(streams))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_kind": "of " + peak_kind, "traffic": None, "kernel": "compose_frames_kernel",
                         "algorithmic_bytes_per_launch": int(b_alg), "bytes_per_mb": b_alg / total_mbs,
                         "kernel_ms": ms / args.steps},
        }
        line["raster"] = {"kernel": "raster_regions_kernel", "ms": raster_ms, "algorithmic_bytes": int(raster_bytes),
                          "achieved": raster_bytes / (raster_ms * 1e-3) / 1e9, "unit": "GB/s",
                          "frac": raster_bytes / (raster_ms * 1e-3) / 1e9 / peak}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                line["roofline"]["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu:
            po = ge.load_oracle()
            cb, _, _ = run_cpu(po, 1, target_seconds=12.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------ second workload: slice rewrite

def main_rewrite(args):
    """batched I-slice rewrite (new header + bit-shifted copy + emulation prevention).  Not the
    north-star metric: `python bench.py --workload rewrite` prints its own line (bytes of source RBSP rewritten per s)."""
    ge.load_package()
    from h264_scroll_encoder_b200 import binding, abi
    gc = binding.GpuComposer(0)
    L = gc.lib
    n, S = args.slices, args.slice_bytes
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, n * S, dtype=np.uint8)
    src[rng.integers(0, len(src), len(src) // 64)] = 0
    idx = rng.integers(0, len(src) - 1, len(src) // 256)
    src[idx] = 0; src[idx + 1] = 0
    hdr = "1000100010000" + "1" + "011" + "010"
    cap = (S * 3 // 2 + 256 + 15) & ~15
    desc = np.zeros(n, abi.REWRITE_DESC)
    desc["hdr_bits"] = abi.hdr_words(hdr)
    desc["src_off"] = np.arange(n, dtype=np.uint64) * np.uint64(S)
    desc["src_bytes"], desc["src_start_bit"], desc["n_bits"] = S, 23, 8 * S - 23
    desc["out_off"] = np.arange(n, dtype=np.uint64) * np.uint64(cap)
    desc["out_cap"], desc["hdr_nbits"], desc["nal_header"] = cap, len(hdr), 0x65
    cb = int(L.h264gpu_rewrite_chunk_bytes())
    max_chunks = (S + 64 + cb - 1) // cb
    d_desc, d_src = gc.to_device(desc), gc.to_device(np.concatenate([src, np.zeros(16, np.uint8)]))
    d_out, d_sizes, d_cnt = gc.alloc(cap * n), gc.alloc(4 * n), gc.alloc(4 * (1 + n * max_chunks))
    ev0, ev1 = L.h264gpu_event_create(), L.h264gpu_event_create()

    def step():
        gc.check(L.h264gpu_rewrite_slices(d_desc.ptr, n, max_chunks, d_src.ptr, d_out.ptr, d_sizes.ptr, d_cnt.ptr, None))
    for _ in range(max(args.warmup, 3)):
        step()
    gc.sync()
    l0 = gc.launch_count()
    gc.check(L.h264gpu_event_record(ev0, None))
    for _ in range(args.steps):
        step()
    gc.check(L.h264gpu_event_record(ev1, None))
    ms = L.h264gpu_event_elapsed_ms(ev0, ev1) / args.steps
    launches = gc.launch_count() - l0
    sizes = d_sizes.download(np.uint32)
    assert (sizes <= cap).all()
    out_bytes = int(sizes.sum())
    # end to end: pinned source up, NAL units back
    p_src = L.h264gpu_host_alloc(len(src)); np.ctypeslib.as_array((ctypes.c_uint8 * len(src)).from_address(p_src))[:] = src
    p_out = L.h264gpu_host_alloc(cap * n)
    def e2e():
        gc.check(L.h264gpu_memcpy_h2d(d_src.ptr, p_src, len(src), None))
        step()
        gc.check(L.h264gpu_memcpy_d2h(p_out, d_out.ptr, cap * n, None))
        gc.sync()
    e2e(); t0 = time.perf_counter()
    for _ in range(3):
        e2e()
    e2e_s = (time.perf_counter() - t0) / 3
    peak, peak_kind = peaks()
    achieved = (n * S + out_bytes) / (ms * 1e-3) / 1e9
    line = {"metric": "rewritten_rbsp_bytes_per_second", "value": n * S / (ms * 1e-3), "unit": "B/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "gpu_launches": int(launches),
            "config": {"workload": "slice rewrite (h264_rewrite_idr_frame on the device): %d I slices of %d bytes per step "
                                   "(1280x720 I_PCM atlas size), %.0f MB in + %.0f MB out vs 126 MB L2"
                                   % (n, S, n * S / 1e6, out_bytes / 1e6)},
            "e2e": {"value": n * S / e2e_s, "unit": "B/s", "h2d_bytes_per_step": int(len(src)), "d2h_bytes_per_step": int(cap * n),
                    "ms_per_step": 1e3 * e2e_s},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_kind": "of " + peak_kind, "traffic": None, "kernel": "rewrite_slices_kernel",
                         "algorithmic_bytes_per_launch": int(n * S + out_bytes)}}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_rewrite_baseline(src[:S].copy(), S)
    print(json.dumps(line))
    return 0


def cpu_rewrite_baseline(rbsp, S):
    """The reference's own h264_rewrite_idr_frame (oracle/_ref/libref_composer.so) on one host core; falls back to this
    library's host C implementation when the reference build is absent."""
    from h264_scroll_encoder_b200 import binding
    path = os.path.join(ROOT, "oracle", "_ref", "libref_composer.so")
    kind = "reference"
    if os.path.exists(path):
        R = ctypes.CDLL(path)
    else:
        R, kind = binding.load_library(), "port"
    cfg, pcfg, nw = binding.ComposerConfig(), binding.ComposerConfig(), binding.NALWriter()
    R.composer_config_init(ctypes.byref(cfg), 1280, 720); R.composer_config_init(ctypes.byref(pcfg), 1280, 720)
    hdr = "1" + "0001000" + "1" + "0000" + "1" + "0" + "0" + "1" + ("010" if cfg.deblocking_filter_control_present_flag else "")
    bits = np.concatenate([np.array([int(c) for c in hdr], np.uint8), np.unpackbits(rbsp)])[:8 * S]
    rb = np.packbits(bits)
    out, tmp = np.zeros(2 * S + 64, np.uint8), np.zeros(2 * S + 64, np.uint8)
    R.h264_rewrite_idr_frame.restype = ctypes.c_size_t
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 5.0:
        R.nal_writer_init(ctypes.byref(nw), ctypes.c_void_p(out.ctypes.data), ctypes.c_size_t(out.size),
                          ctypes.c_void_p(tmp.ctypes.data), ctypes.c_size_t(tmp.size))
        assert R.h264_rewrite_idr_frame(ctypes.byref(nw), ctypes.byref(cfg), ctypes.byref(pcfg), ctypes.c_void_p(rb.ctypes.data),
                                        ctypes.c_size_t(len(rb))) > 0
        reps += 1
    wall = time.perf_counter() - t0
    return {"value": reps * S / wall, "unit": "B/s", "cores": 1, "kind": kind,
            "sample": "%d calls of h264_rewrite_idr_frame on one %d-byte slice, %.1f s wall" % (reps, S, wall)}


def main_unescape(args):
    """batched EBSP -> RBSP (ebsp_to_rbsp on the device).  `python bench.py --workload unescape`."""
    ge.load_package()
    from h264_scroll_encoder_b200 import binding, abi
    gc = binding.GpuComposer(0)
    L = gc.lib
    n, S = args.slices, (args.slice_bytes + 15) & ~15
    rng = np.random.default_rng(9)
    src = rng.integers(0, 256, n * S, dtype=np.uint8)
    idx = rng.integers(0, len(src) - 4, len(src) // 512)
    src[idx] = 0; src[idx + 1] = 0; src[idx + 2] = 3; src[idx + 3] = 1          # an emulation prevention byte every ~512 bytes
    desc = np.zeros(n, abi.UNESCAPE_DESC)
    desc["src_off"] = np.arange(n, dtype=np.uint64) * np.uint64(S)
    desc["src_bytes"], desc["out_cap"] = S, S
    desc["out_off"] = desc["src_off"]
    cb = int(L.h264gpu_unescape_chunk_bytes())
    max_chunks = (S + cb - 1) // cb
    d_desc, d_src = gc.to_device(desc), gc.to_device(np.concatenate([src, np.zeros(32, np.uint8)]))
    d_out, d_sizes, d_st = gc.alloc(n * S + 32), gc.alloc(4 * n), gc.alloc(4 * (1 + n * max_chunks))
    ev0, ev1 = L.h264gpu_event_create(), L.h264gpu_event_create()

    def step():
        gc.check(L.h264gpu_unescape_nals(d_desc.ptr, n, max_chunks, d_src.ptr, d_out.ptr, d_sizes.ptr, d_st.ptr, None))
    for _ in range(max(args.warmup, 3)):
        step()
    gc.sync()
    l0 = gc.launch_count()
    gc.check(L.h264gpu_event_record(ev0, None))
    for _ in range(args.steps):
        step()
    gc.check(L.h264gpu_event_record(ev1, None))
    ms = L.h264gpu_event_elapsed_ms(ev0, ev1) / args.steps
    launches = gc.launch_count() - l0
    out_bytes = int(d_sizes.download(np.uint32).sum())
    peak, peak_kind = peaks()
    achieved = (n * S + out_bytes) / (ms * 1e-3) / 1e9
    line = {"metric": "unescaped_ebsp_bytes_per_second", "value": n * S / (ms * 1e-3), "unit": "B/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "gpu_launches": int(launches),
            "config": {"workload": "EBSP -> RBSP (ebsp_to_rbsp on the device): %d payloads of %d bytes per step, %.0f MB in + %.0f MB out "
                                   "vs 126 MB L2" % (n, S, n * S / 1e6, out_bytes / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_kind": "of " + peak_kind, "traffic": None, "kernel": "unescape_kernel",
                         "algorithmic_bytes_per_launch": int(n * S + out_bytes)}}
    if not args.no_cpu:
        path = os.path.join(ROOT, "oracle", "_ref", "libref_composer.so")
        R, kind = (ctypes.CDLL(path), "reference") if os.path.exists(path) else (binding.load_library(), "port")
        R.ebsp_to_rbsp.restype = ctypes.c_size_t
        one, out = src[:S].copy(), np.zeros(S + 8, np.uint8)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 5.0:
            R.ebsp_to_rbsp(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(one.ctypes.data), ctypes.c_size_t(S))
            reps += 1
        wall = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": reps * S / wall, "unit": "B/s", "cores": 1, "kind": kind,
                                "sample": "%d calls of ebsp_to_rbsp on one %d-byte payload, %.1f s wall" % (reps, S, wall)}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=128, help="streams per GPU")
    ap.add_argument("--frames", type=int, default=64, help="frames per stream per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="compose", choices=["compose", "rewrite", "unescape"],
                    help="compose = the north-star metric (default); rewrite = batched I-slice rewrite (SURVEY 8f)")
    ap.add_argument("--slices", type=int, default=128)
    ap.add_argument("--slice-bytes", type=int, default=1382400)
    ap.add_argument("--e2e-zero-copy", action="store_true",
                    help="end-to-end path whose pack kernel writes into pinned host memory instead of a separate D2H copy")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not move the process to the GPU's NUMA node")
    ap.add_argument("--e2e-python", action="store_true", help="drive the end-to-end pipeline from Python (chunk by chunk) instead of the single C-ABI call")
    ap.add_argument("--e2e-streams", type=int, default=4)
    ap.add_argument("--e2e-lag", type=int, default=2, help="chunks in flight before the first packed chunk is fetched")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="frame chunks the end-to-end path pipelines over streams")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.workload == "rewrite":
        return main_rewrite(args)
    if args.workload == "unescape":
        return main_unescape(args)
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
