#!/usr/bin/env python
"""
bench.py - composed macroblocks/second of bit-exact P-frames (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a, through the C ABI)
  python bench.py --impl reference ...                     the reference's own CPU code on the host cores

Workload (config.workload): BASELINE.json configs[4] "batched 1024 independent 1080p UI streams sharded
across 8xB200": every GPU takes 128 streams x 64 frames of the config-3 layout (static chrome, three
carousels, a two-reference vertical grid, a 25x25-MB spliced dynamic region), so the
8-GPU run is exactly the 1024-stream batch and per-GPU work is fixed (weak scaling, no collective).
--workload config4 | config2 | config1 time the other named resolutions the same way (each line with its own
roofline, cpu_baseline and e2e); the default line carries their kernel-only fractions under "other_configs",
the side kernels' under "side_kernels" and the per-call latency of the drop-in frame writer under "latency".

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


class _HintsView:
    """One frame's hints behaving like the whole batch's array for tests/helpers.py (which slices [f * stride : ...])."""

    def __init__(self, frame_hints, first):
        self.h, self.first = frame_hints, first

    def __getitem__(self, sl):
        return self.h[sl.start - self.first:sl.stop - self.first]


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~0.5 ms during the timed region
    (the region lasts tens of milliseconds: too short for nvidia-smi's 100 ms loop)."""

    def __init__(self, index):
        self.index, self.sm, self.ts, self.reasons, self.stop_flag, self.t, self.max = index, [], [], 0, False, None, None
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
                self.ts.append(time.perf_counter())
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
    """Run this process on the CPUs next to GPU `index`, BEFORE any pinned memory is allocated (pinned pages are placed
    where the allocating thread runs): the end-to-end path moves ~0.5 GB per step per GPU through pinned host buffers,
    and with several ranks on a multi-socket host buffers on the wrong node cross the socket link twice.
    Sources, in order: the GPU's numa_node in sysfs; NVML's CPU affinity for the device (what `nvidia-smi topo` prints)
    when sysfs says -1 (virtualised hosts); a host with one NUMA node needs no binding.
    Returns what was done: {"node": n | None, "cpus": count, "how": "..."}."""
    def cpus_of_node(node):
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        nodes = []
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
        node = -1
        try:
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        except (OSError, ValueError):
            pass
        allowed = os.sched_getaffinity(0)
        if node >= 0:
            cpus = cpus_of_node(node) & allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
                return {"node": node, "cpus": len(cpus), "how": "sysfs numa_node of the GPU"}
        if len(nodes) <= 1:
            return {"node": nodes[0] if nodes else 0, "cpus": len(allowed), "how": "host has a single NUMA node: nothing to bind"}
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + i for w, m in enumerate(mask) for i in range(64) if (int(m) >> i) & 1} & allowed
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            on = [n for n in nodes if cpus_of_node(n) & cpus]
            return {"node": on[0] if len(on) == 1 else None, "cpus": len(cpus), "how": "NVML CPU affinity of the GPU (sysfs numa_node is -1)"}
        return {"node": None, "cpus": len(allowed), "how": "no topology information (sysfs numa_node -1, NVML affinity = all CPUs)"}
    except Exception as e:                                        # noqa: BLE001
        return {"node": None, "cpus": len(os.sched_getaffinity(0)), "how": "not bound: %r" % (e,)}


# --------------------------------------------------------------------------------------- reference arm

def cpu_sample(po, args):
    """Bounded sample of the selected workload as dense hints for the CPU code: (batch, hints[n, mbs], tables[n, stride] | None,
    payload | None, what)."""
    from h264_scroll_encoder_b200 import workloads, abi
    wl = args.workload
    if wl in ("compose", "config5spec"):                    # (the reference has one mode: skip OFF, its own predictor)
        b, what = workloads.make_config5(2, 8), "2 streams x 8 frames of the config-5 workload"
    elif wl == "config4":
        b, what = workloads.make_config4(4), "4 frames of config 4"
    elif wl == "config2":
        b, what = workloads.make_config2(16), "16 frames of config 2"
    else:
        b, what = workloads.make_config1(1, 32), "32 frames of config 1 (coded with their own num_refs; one header for all)"
    hints = np.zeros((b.n_frames, b.n_mbs), abi.HINT)
    for f in range(b.n_frames):
        r0, rc = int(b.rdesc["region_first"][f]), int(b.rdesc["region_count"][f])
        di = b.dyn_info[int(b.rdesc["dyn_info_off"][f]):] if b.dyn is not None else None
        hints[f] = po.raster_regions_numpy(b.regions[r0:r0 + rc], b.mb_w, b.mb_h, di)
    tab = None
    if b.dyn is not None:                                    # per-frame tables re-based on one shared payload buffer
        stride = int(b.dyn["n_mbs"][0]) + 1
        tab = np.zeros((b.n_frames, stride), np.uint32)
        for f in range(b.n_frames):
            t0 = int(b.dyn["table_off"][f])
            tab[f] = b.dyn_tab[t0:t0 + stride] + np.uint32(8 * int(b.dyn["payload_off"][f]))
        tab = np.ascontiguousarray(tab)
    return b, np.ascontiguousarray(hints), tab, b.payload, what


def run_cpu(po, threads, target_seconds, args, kind=None):
    """Time the reference's CPU path (oracle/_ref when present, else the oracle port) on `threads` threads."""
    use_ref = po.have_ref() if kind is None else (kind == "reference")
    chk = po.Ref() if use_ref else po.Oracle()
    b, hints, tab, payload, what = cpu_sample(po, args)
    hdr = b.desc["hdr_bits"][b.n_frames - 1].copy()
    nb = int(b.desc["hdr_nbits"][b.n_frames - 1])
    nrefs = int(b.desc["num_refs"][b.n_frames - 1])
    stride = tab.shape[1] if tab is not None else 0
    # calibrate reps on one thread
    t1, _ = chk.time_hinted(hints, b.n_frames, b.mb_w, b.mb_h, nrefs, 0, hdr, nb, payload, tab, stride, 1)
    reps = max(1, int(target_seconds / max(t1, 1e-6)))
    res = [None] * threads

    def work(i):
        res[i] = chk.time_hinted(hints, b.n_frames, b.mb_w, b.mb_h, nrefs, 0, hdr, nb, payload, tab, stride, reps)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    mbs = threads * reps * b.n_frames * b.n_mbs
    return {"value": mbs / wall, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": "%d threads x %d passes over %d frames (%s, %d MBs/frame, dense hints), %.1f s wall"
                      % (threads, reps, b.n_frames, what, b.n_mbs, wall)}, wall, reps


def main_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    po = ge.load_oracle()
    ge.load_package()
    threads = os.cpu_count() or 1
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        r, wall, reps = run_cpu(po, threads, max(1.0, 20.0 / max(args.steps + args.warmup, 1)), args)
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


def build_batch(args, rank=0, world=1):
    """The rank's batch of the selected workload (region lists + payload; every frame an independent unit of work)."""
    from h264_scroll_encoder_b200 import workloads, sharding
    wl = args.workload
    if wl == "compose":                                       # config 5: weak scaling, args.streams streams per rank
        first, count = sharding.shard_streams(args.streams * world, rank, world)
        return workloads.make_config5(count, args.frames, stream0=first)
    if wl == "config4":                                       # 64 frames (t = 0..63) x copies: 614 MB of hints at 37 copies
        return workloads.repeat_batch(workloads.make_config4(64), args.copies or DEFAULT_COPIES[wl])
    if wl == "config2":                                       # 250 frames x copies: 720p + splice
        return workloads.repeat_batch(workloads.make_config2(250), args.copies or DEFAULT_COPIES[wl])
    if wl == "config1":                                       # 248 frames x streams
        return workloads.make_config1(args.copies or DEFAULT_COPIES[wl], 248)
    if wl == "config5spec":                                   # config-5 streams coded with spec P_Skip + the spec median predictor
        from h264_scroll_encoder_b200 import abi
        return workloads.make_config5(args.copies or DEFAULT_COPIES[wl], 64, flags=abi.SKIP_SPEC | abi.PRED_SPEC)
    raise SystemExit("unknown workload " + wl)


# Copies of the other named configurations per GPU and step: larger than L2, and a whole number of frames per resident CTA
# of the persistent compose grid (148 SMs x 2 CTAs, one frame per CTA at a time) - 64 x 37 = 8 x 296, 248 x 37 = 31 x 296,
# 250 x 13 = 10.98 x 296 - so that the figure is the kernel's rate, not the last partial wave of a small batch
# (64 x 8 frames of config 4 were 1.73 waves: 0.170 of the roofline against the 0.206 the same kernel does on 8 whole waves).
DEFAULT_COPIES = {"config4": 37, "config2": 13, "config1": 37, "config5spec": 37}


def workload_config(args, b=None):
    names = {
        "compose": "config5: batched independent 1080p UI streams, %d streams x %d frames per GPU (1920x1088 = 120x68 MBs, "
                   "config-3 layout with a 25x25-MB spliced dynamic region, num_refs 2, skip OFF, reference predictor)"
                   % (args.streams, args.frames),
        "config4": "config4: 3840x2160 (240x135 MBs) full-frame scroll with per-row varying MVs, every MB a non-zero mvd, "
                   "64 frames x %d copies per GPU, num_refs 2, skip OFF, reference predictor" % (args.copies or DEFAULT_COPIES["config4"]),
        "config2": "config2: 1280x720 legacy scroll 0->720->0 at speed 4 with a 25x25-MB spliced dynamic region, 250 frames x "
                   "%d copies per GPU, num_refs 2, skip OFF, reference predictor" % (args.copies or DEFAULT_COPIES["config2"]),
        "config1": "config1: experiments/scroll-encoder twin, 1280x720, 248 frames from offset 496 at speed 1 (first frame the "
                   "waypoint frame, num_refs 2 -> 3) x %d streams per GPU, skip OFF, reference predictor" % (args.copies or DEFAULT_COPIES["config1"]),
    }
    names["config5spec"] = ("config5 streams in the SPEC modes (SURVEY 8f rank 3: mb_skip_run / P_Skip and the median predictor as H.264 "
                            "8.4.1.1-8.4.1.3 have them - no counterpart in the reference; checked against the oracle port): %d streams x 64 "
                            "frames of 1920x1088 per GPU with the 25x25-MB spliced region, num_refs 2" % (args.copies or DEFAULT_COPIES["config5spec"]))
    shape = {"compose": (args.streams * args.frames, 8160), "config4": (64 * (args.copies or DEFAULT_COPIES["config4"]), 32400),
             "config5spec": (64 * (args.copies or DEFAULT_COPIES["config5spec"]), 8160),
             "config2": (250 * (args.copies or DEFAULT_COPIES["config2"]), 3600),
             "config1": (248 * (args.copies or DEFAULT_COPIES["config1"]), 3600)}[args.workload]
    cfg = {"workload": names[args.workload], "sharding": "streams block-partitioned over ranks, no collective",
           "frames_per_gpu": shape[0], "mbs_per_frame": shape[1],
           "l2_policy": "inputs larger than L2: %.0f MB of dense hints (+ the spliced payload) per GPU per step vs 126 MB L2"
                        % (8 * shape[0] * shape[1] / 1e6)}
    if args.workload == "compose":
        cfg.update({"streams_per_gpu": args.streams, "frames_per_stream": args.frames})
    if b is not None:
        assert (b.n_frames, b.n_mbs) == shape
    return cfg


def algorithmic_bytes(b, sizes):
    """hints once (8 B/MB) + payload once + offset tables + output once + descriptors."""
    n = b.n_frames
    s_dyn = tab = 0
    if b.dyn is not None:
        ends = b.dyn["table_off"].astype(np.int64) + b.dyn["n_mbs"].astype(np.int64)
        s_dyn = int((b.dyn_tab[ends] // 8).sum())
        tab = int(4 * (b.dyn["n_mbs"].astype(np.int64) + 1).sum())
    return 8 * n * b.n_mbs + s_dyn + tab + int(sizes.sum()) + 64 * n


def check_bit_exact(po, b, hints_of, frames_of, n_check=64, seed=11):
    """>= n_check frames of the timed run's output against the reference's own functions (oracle/_ref) or the port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    spec = bool((np.asarray(b.desc["flags"]) != 0).any())        # SPEC modes: the reference has none - the oracle port checks
    chk, kind = (po.Ref(), "reference") if po.have_ref() and not spec else (po.Oracle(), "port")
    pick = np.random.default_rng(seed).choice(b.n_frames, min(n_check, b.n_frames), replace=False)
    for f in pick:
        f = int(f)
        want = helpers.checker_frame(chk, b, hints_of(f), f)
        got = frames_of(f)
        if not np.array_equal(got, want):
            return {"bit_exact": False, "frames_checked": int(len(pick)), "checker": kind, "first_bad_frame": f}
    return {"bit_exact": True, "frames_checked": int(len(pick)), "checker": kind}


# --------------------------------------------------------------------------------------------- our arm

def main_ours(args):
    rank, local, world = dist_env()
    ge.load_package()
    from h264_scroll_encoder_b200 import binding, workloads, abi
    dist = None
    if world > 1:
        # the data path has no collective (nothing reduces across streams): the process group only carries the barrier
        # around the timed region and the max / sum of the timings - gloo, no NCCL anywhere
        import torch.distributed as dist
        dist.init_process_group("gloo")
    numa = {"node": None, "cpus": len(os.sched_getaffinity(0)), "how": "--no-numa-bind"} if args.no_numa_bind else bind_to_gpu_numa_node(local)
    numa_node = numa["node"]
    gc = binding.GpuComposer(local)      # raises if the CUDA library or a B200 is missing: no fallback
    L = gc.lib

    b = build_batch(args, rank, world)
    n, n_mbs = b.n_frames, b.n_mbs
    has_dyn = b.dyn is not None
    total_mbs = n * n_mbs

    # ---- device-resident inputs -------------------------------------------------------------------
    d_desc, d_rdesc, d_reg = gc.to_device(b.desc), gc.to_device(b.rdesc), gc.to_device(b.regions)
    d_info = d_dyn = d_tab = d_pay = None
    if has_dyn:
        pay_pad = np.concatenate([b.payload, np.zeros(16, np.uint8)])
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

    sampler = ClockSampler(local)                 # (started in front of the warm-up: its first NVML reads take milliseconds)
    sampler.start()
    for _ in range(args.warmup):
        kernel_step()
    barrier()
    launches0 = gc.launch_count()
    t_region0 = time.perf_counter()
    gc.check(L.h264gpu_event_record(ev0, None))
    for _ in range(args.steps):
        kernel_step()
    gc.check(L.h264gpu_event_record(ev1, None))
    ms = L.h264gpu_event_elapsed_ms(ev0, ev1)
    t_region1 = time.perf_counter()
    launches = gc.launch_count() - launches0
    samples_timed = sum(1 for t in list(sampler.ts) if t_region0 <= t <= t_region1)
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
    raster_bytes = 8 * total_mbs + b.regions.nbytes + b.rdesc.nbytes + (b.dyn_info.nbytes if has_dyn else 0)   # map written, lists and dyn-info read

    if hasattr(L, "h264gpu_debug_trace"):                         # event-trace experiment build only
        tr = (ctypes.c_uint64 * (32 * 1024))()
        L.h264gpu_debug_trace(tr)
        os.makedirs("gpurun_out", exist_ok=True)
        np.frombuffer(tr, np.uint64).tofile("gpurun_out/trace.bin")
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
    b_alg = algorithmic_bytes(b, sizes)

    # ---- the metric says "bit-exact": sampled frames of the timed run against the reference's functions ----
    exact = None
    if rank == 0 and not ablate and not args.no_check:
        po_ = ge.load_oracle()
        exact = check_bit_exact(
            po_, b,
            lambda f: _HintsView(d_hints.download(abi.HINT, count=b.n_mbs, offset=8 * b.hint_stride * f), b.hint_stride * f),
            lambda f: d_out.download(np.uint8, count=int(sizes[f]), offset=int(b.desc["out_off"][f])))
        assert exact["bit_exact"], "GPU output differs from the %s checker in frame %d" % (exact["checker"], exact["first_bad_frame"])

    # ---- end to end: host buffers in, packed Annex-B stream out, every step ------------------------
    # ONE C-ABI call (h264gpu_compose_regions_host, the call a host application makes): pinned host region lists +
    # payload in, packed Annex-B stream out.  The library cuts the batch into frame chunks and keeps upload, kernels and
    # download on three streams (PCIe is full duplex); every input byte crosses the bus every step, every stream byte
    # comes back.
    n_chunks = max(1, min(args.e2e_chunks, n))
    hb, hb_keep = gc.make_host_batch(b, pinned=True)
    h2d_bytes = b.h2d_bytes()
    h_stream_cap = b.slot_cap * n // 2 if has_dyn else b.slot_cap * n
    h_out = L.h264gpu_host_alloc(h_stream_cap)
    h_offs_all = L.h264gpu_host_alloc(8 * (n + 1))
    h_sizes_all = L.h264gpu_host_alloc(4 * n)
    offs_all_view = np.ctypeslib.as_array((ctypes.c_uint64 * (n + 1)).from_address(h_offs_all))

    def e2e_step():
        gc.compose_regions_host(hb, h_out, h_stream_cap, h_offs_all, h_sizes_all, n_chunks)
        return int(offs_all_view[n])

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
    if not ablate:
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
        from h264_scroll_encoder_b200 import sharding
        (ms_step, e2e_step_s), (job_mbs, _) = sharding.reduce_job(dist, "cpu", [ms_step, e2e_step_s], [float(total_mbs), float(b_alg)])
    else:
        job_mbs = float(total_mbs)

    if rank == 0:
        peak, peak_kind = peaks()
        achieved = b_alg / (ms / args.steps * 1e-3) / 1e9      # this rank's kernel, GB/s
        line = {
            "metric": METRIC, "value": job_mbs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, b), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": job_mbs / e2e_step_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h + 8 * (n + n_chunks) + 4 * n), "ms_per_step": 1e3 * e2e_step_s,
                    "chunks": n_chunks, "gpu_launches_per_step": e2e_launches, "host_numa_node": numa_node, "host_binding": numa,
                    "path": "one h264gpu_compose_regions_host call: pinned host region lists + payload -> H2D -> raster -> compose "
                            "-> pack -> D2H packed stream, in %d frame chunks over 3 streams" % n_chunks},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_kind": "of " + peak_kind, "traffic": None, "kernel": "compose_frames_kernel",
                         "algorithmic_bytes_per_launch": int(b_alg), "bytes_per_mb": b_alg / total_mbs,
                         "kernel_ms": ms / args.steps},
        }
        if exact is not None:
            line.update(exact)
        # the end-to-end number against what this pool's hosts can move with `world` ranks copying both ways at once
        # (profiles/pcie_scaling.json, measured by scripts/pcie_scaling.py on one 8 x B200 box)
        try:
            tab = json.load(open(os.path.join(ROOT, "profiles", "pcie_scaling.json")))["ranks"].get(str(world))
            if tab:
                bw = tab["both"]["per_rank_mean_gbs"]
                floor_ms = max(h2d_bytes, d2h) / (bw * 1e9) * 1e3
                line["e2e"]["bus_ceiling"] = {"both_ways_gbs_per_rank": bw, "floor_ms_per_step": floor_ms,
                                              "floor_over_measured": floor_ms / (1e3 * e2e_step_s),
                                              "source": "profiles/pcie_scaling.json (concurrent pinned copies, %d ranks)" % world}
        except Exception:                                         # noqa: BLE001
            pass
        line["raster"] = {"kernel": "raster_regions_kernel", "ms": raster_ms, "algorithmic_bytes": int(raster_bytes),
                          "achieved": raster_bytes / (raster_ms * 1e-3) / 1e9, "unit": "GB/s",
                          "frac": raster_bytes / (raster_ms * 1e-3) / 1e9 / peak}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:                                                  # (one ncu capture of this workload's full-size launch, not measured in this run)
                t = json.load(open(tr))
                t = t if args.workload == "compose" else t.get("other_workloads", {}).get(args.workload, {})
                same_batch = args.workload != "compose" or (args.streams, args.frames) == (128, 64)
                if same_batch and t.get("dram_bytes_per_launch"):
                    line["roofline"]["traffic"] = t["dram_bytes_per_launch"]
                    line["roofline"]["traffic_source"] = "profiles/traffic.json (ncu capture of this workload's launch; not measured in this run)"
            except Exception:
                pass
        if world == 1 and not args.no_cpu:
            po = ge.load_oracle()
            cb, _, _ = run_cpu(po, 1, 10.0, args)
            line["cpu_baseline"] = cb
        if world == 1 and args.workload == "compose" and not args.no_extras:
            line.update(default_line_extras(args, gc))
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def quick_kernel_fraction(gc, args, wl):
    """Kernel-only timing of another named configuration (inputs resident in HBM, larger than L2): its roofline fraction."""
    from h264_scroll_encoder_b200 import abi
    a = argparse.Namespace(**vars(args))
    a.workload, a.copies = wl, 0
    b = build_batch(a)
    L, n = gc.lib, b.n_frames
    d_desc, d_rdesc, d_reg = gc.to_device(b.desc), gc.to_device(b.rdesc), gc.to_device(b.regions)
    d_info = d_dyn = d_tab = d_pay = None
    if b.dyn is not None:
        d_info, d_dyn, d_tab = gc.to_device(b.dyn_info), gc.to_device(b.dyn), gc.to_device(b.dyn_tab)
        d_pay = gc.to_device(np.concatenate([b.payload, np.zeros(16, np.uint8)]))
    d_hints, d_out, d_sizes = gc.alloc(8 * b.hint_stride * n), gc.alloc(b.slot_cap * n), gc.alloc(4 * n)
    gc.raster_regions(d_rdesc, n, b.n_mbs, d_reg, d_info, d_hints)
    ev0, ev1 = L.h264gpu_event_create(), L.h264gpu_event_create()
    for _ in range(3):
        gc.compose_batch(d_desc, n, d_hints, d_dyn, d_tab, d_pay, d_out, d_sizes)
    gc.sync()
    steps = 5
    gc.check(L.h264gpu_event_record(ev0, None))
    for _ in range(steps):
        gc.compose_batch(d_desc, n, d_hints, d_dyn, d_tab, d_pay, d_out, d_sizes)
    gc.check(L.h264gpu_event_record(ev1, None))
    ms = L.h264gpu_event_elapsed_ms(ev0, ev1) / steps
    sizes = d_sizes.download(np.uint32)
    assert (sizes <= b.slot_cap).all() and (sizes > 5).all()
    exact = check_bit_exact(ge.load_oracle(), b,
                            lambda f: _HintsView(d_hints.download(abi.HINT, count=b.n_mbs, offset=8 * b.hint_stride * f), b.hint_stride * f),
                            lambda f: d_out.download(np.uint8, count=int(sizes[f]), offset=int(b.desc["out_off"][f])), n_check=16)
    b_alg = algorithmic_bytes(b, sizes)
    peak, _ = peaks()
    for x in (d_desc, d_rdesc, d_reg, d_info, d_dyn, d_tab, d_pay, d_hints, d_out, d_sizes):
        if x is not None:
            x.free()
    return {"workload": workload_config(a, b)["workload"], "kernel_ms": ms, "value": n * b.n_mbs / (ms * 1e-3), "unit": UNIT,
            "bytes_per_mb": b_alg / (n * b.n_mbs), "achieved": b_alg / (ms * 1e-3) / 1e9, "frac": b_alg / (ms * 1e-3) / 1e9 / peak,
            "bit_exact": exact["bit_exact"], "frames_checked": exact["frames_checked"]}


def default_line_extras(args, gc):
    """What the default line carries beside the config-5 numbers, so that the driver records them too."""
    extras = {"other_configs": {}, "side_kernels": {}}
    for wl in ("config4", "config2", "config1", "config5spec"):
        extras["other_configs"][wl] = quick_kernel_fraction(gc, args, wl)
    a = argparse.Namespace(**vars(args))
    a.no_cpu, a.steps, a.warmup = True, 10, 3
    for name, fn in (("rewrite", main_rewrite), ("unescape", main_unescape)):
        ln = fn(a)
        extras["side_kernels"][name] = {"kernel": ln["roofline"]["kernel"], "ms": ln["ms_per_step"], "frac": ln["roofline"]["frac"],
                                        "achieved": ln["roofline"]["achieved"], "unit": "GB/s", "workload": ln["config"]["workload"]}
    a.no_cpu = args.no_cpu
    extras["latency"] = main_latency(a)["resolutions"]
    return extras


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
    return line


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
    return line


def main_latency(args):
    """Per-call latency of the drop-in frame writer: h264_write_scroll_p_frame (one frame per call, the reference's API,
    src/composer.c:255-264) through this library on the GPU, and the reference's own function on one host core."""
    ge.load_package()
    from h264_scroll_encoder_b200 import binding, workloads
    gc = binding.GpuComposer(0)
    lib = gc.lib
    lib.h264_write_scroll_p_frame.restype = ctypes.c_size_t
    out = {}
    po = ge.load_oracle()
    for (w, h) in ((1280, 720), (1920, 1088)):
        offs = [int(o) for o in workloads.legacy_schedule(400, 1, 480)]      # below the first waypoint: scroll frames only
        cfg = binding.ComposerConfig()
        lib.composer_config_init(ctypes.byref(cfg), w, h)
        lib.composer_config_set_sps_params(ctypes.byref(cfg), 4, 2, 4)
        lib.composer_config_set_pps_params(ctypes.byref(cfg), 1, 1)
        cfg.frame_num = 2
        buf = np.zeros(8 << 20, np.uint8)
        nw = binding.NALWriter()
        lat = []
        for i, o in enumerate(offs):
            if i % 100 == 0:
                lib.nal_writer_init(ctypes.byref(nw), ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.size), None, ctypes.c_size_t(0))
            t0 = time.perf_counter()
            n = lib.h264_write_scroll_p_frame(ctypes.byref(nw), ctypes.byref(cfg), o)
            lat.append(time.perf_counter() - t0)
            assert n > 0
        lat = np.array(lat[50:]) * 1e6
        e = {"us_per_call_median": float(np.median(lat)), "us_per_call_p90": float(np.percentile(lat, 90)), "calls": int(len(lat))}
        if po.have_ref() and not args.no_cpu:
            t, nbytes, nals = po.Ref().time_legacy(w, h, np.array(offs, np.int32), reps=5)
            e["reference_us_per_call"] = 1e6 * t / nals
            e["reference_kind"] = "reference (h264_write_scroll_p_frame of /root/reference, one host core)"
        out["%dx%d" % (w, h)] = e
    return {"metric": "frame_writer_latency", "unit": "us per h264_write_scroll_p_frame call", "higher_is_better": False,
            "path": "pinned descriptors read by the kernels over PCIe -> raster -> compose writes the NAL unit and its size into pinned "
                    "host memory -> one stream sync -> memcpy into the NALWriter", "resolutions": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=128, help="streams per GPU")
    ap.add_argument("--frames", type=int, default=64, help="frames per stream per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="compose", choices=["compose", "config4", "config2", "config1", "config5spec", "rewrite", "unescape", "latency"],
                    help="compose = the north-star metric on config 5 (default); config4 / config2 / config1 = the other named "
                         "resolutions; rewrite / unescape = side kernels (SURVEY 8f); latency = per-call latency of the frame writer")
    ap.add_argument("--copies", type=int, default=0, help="config4/2/1: copies of the configuration's frames per GPU (0 = enough to exceed L2)")
    ap.add_argument("--slices", type=int, default=128)
    ap.add_argument("--slice-bytes", type=int, default=1382400)
    ap.add_argument("--no-numa-bind", action="store_true", help="do not move the process to the GPU's NUMA node")
    ap.add_argument("--no-check", action="store_true", help="skip the bit-exact check of sampled frames")
    ap.add_argument("--no-extras", action="store_true", help="default line without other_configs / side_kernels / latency")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="frame chunks the end-to-end path pipelines over streams")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.workload in ("rewrite", "unescape", "latency"):
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the compose workloads only"}))
            return 0
        print(json.dumps({"rewrite": main_rewrite, "unescape": main_unescape, "latency": main_latency}[args.workload](args)))
        return 0
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
