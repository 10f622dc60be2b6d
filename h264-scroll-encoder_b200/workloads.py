"""
Synthetic workloads of SURVEY.md section 8(d) / Appendix B, shared by the GPU bench, the CPU baseline and
the parity tests.  Everything is produced as *region lists* (what a UI layer would hand over) plus the
pre-encoded payload of the dynamic region; dense hint maps are derived from them (on the device by
h264gpu_raster_regions, or by the numpy checker in oracle/pyoracle.py).

All randomness is the splitmix64-style `mix(seed, a, b)` of Appendix B, so every consumer regenerates
identical inputs from (config, stream, frame).
"""
import os
import numpy as np
from . import abi

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix(seed, a, b):
    """Appendix B: splitmix64 finaliser over three words (vectorised, uint64)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + np.asarray(a, dtype=np.uint64) * np.uint64(0xBF58476D1CE4E5B9)
             + np.asarray(b, dtype=np.uint64) * np.uint64(0x94D049BB133111EB))
        z = z ^ (z >> np.uint64(30)); z = z * np.uint64(0xBF58476D1CE4E5B9)
        z = z ^ (z >> np.uint64(27)); z = z * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def legacy_schedule(n, speed, max_off, start=0):
    """0 -> max -> 0 ping-pong (src/main.c:108-120; experiments/scroll-encoder/src/main.c:402-415)."""
    i = np.arange(n, dtype=np.int64)
    cyc = 2 * max_off
    pos = (i * speed + start) % cyc
    return np.where(pos < max_off, pos, cyc - pos).astype(np.int32)


def p_header_bits(frame_num, num_refs=2, log2_max_frame_num=4, deblock=1):
    """31-bit P-slice header of a 2-reference, no-waypoint frame (src/h264_writer.c:455-488)."""
    assert num_refs == 2
    fn = frame_num % (1 << log2_max_frame_num)
    s = "111" + format(fn, "0%db" % log2_max_frame_num) + "1" + "010" + "1" + "011" + "1" + "011" + "010" + "00100" + "1"
    if deblock:
        s += "010"
    return s


class Batch:
    """Host-side description of a batch of frames (numpy arrays laid out as the C ABI expects)."""

    def __init__(self, mb_w, mb_h):
        self.mb_w, self.mb_h = mb_w, mb_h
        self.n_mbs = mb_w * mb_h
        self.hint_stride = (self.n_mbs + 1) & ~1
        self.desc = None            # abi.FRAME_DESC [n]
        self.rdesc = None           # abi.RASTER_DESC [n]
        self.regions = None         # abi.REGION [...]
        self.dyn_info = None        # abi.HINT [...]  (per spliced MB: neighbour-visible motion + dyn index)
        self.dyn = None             # abi.DYN_DESC [n] or None
        self.dyn_tab = None         # uint32 [...]
        self.payload = None         # uint8 [...]
        self.slot_cap = 0
        self.name = ""

    @property
    def n_frames(self):
        return len(self.desc)

    def total_mbs(self):
        return self.n_frames * self.n_mbs

    def h2d_bytes(self):
        n = self.desc.nbytes + self.rdesc.nbytes + self.regions.nbytes
        for a in (self.dyn_info, self.dyn, self.dyn_tab, self.payload):
            if a is not None:
                n += a.nbytes
        return n


def _finish(batch, headers, regions_per_frame, num_refs, flags, slot_cap, dyn_parts=None, nal_headers=None):
    n = len(headers)
    b = batch
    b.slot_cap = int(slot_cap)
    desc = np.zeros(n, abi.FRAME_DESC)
    rdesc = np.zeros(n, abi.RASTER_DESC)
    regs = []
    first = 0
    for f in range(n):
        bits = headers[f]
        desc["hdr_bits"][f] = abi.hdr_words(bits)
        desc["hdr_nbits"][f] = len(bits)
        rdesc["region_first"][f] = first
        rdesc["region_count"][f] = len(regions_per_frame[f])
        regs.append(regions_per_frame[f])
        first += len(regions_per_frame[f])
    desc["hint_off"] = np.arange(n, dtype=np.uint64) * np.uint64(b.hint_stride)
    desc["out_off"] = np.arange(n, dtype=np.uint64) * np.uint64(slot_cap)
    desc["out_cap"] = slot_cap
    desc["dyn_id"] = abi.NO_DYN
    desc["mb_w"], desc["mb_h"] = b.mb_w, b.mb_h
    desc["num_refs"] = num_refs
    desc["flags"] = flags
    desc["nal_header"] = 0x01 if nal_headers is None else nal_headers
    rdesc["hint_off"] = desc["hint_off"]
    rdesc["mb_w"], rdesc["mb_h"] = b.mb_w, b.mb_h
    b.desc, b.rdesc = desc, rdesc
    b.regions = np.concatenate(regs) if regs else np.zeros(0, abi.REGION)
    if dyn_parts is not None:
        infos, tabs, pays = dyn_parts
        dyn = np.zeros(n, abi.DYN_DESC)
        info_off = tab_off = pay_off = 0
        for f in range(n):
            rdesc["dyn_info_off"][f] = info_off
            dyn["table_off"][f] = tab_off
            dyn["payload_off"][f] = pay_off
            dyn["n_mbs"][f] = len(infos[f])
            dyn["payload_bytes"][f] = len(pays[f])
            desc["dyn_id"][f] = f
            info_off += len(infos[f]); tab_off += len(tabs[f]); pay_off += len(pays[f])
        b.dyn = dyn
        b.dyn_info = np.concatenate(infos)
        b.dyn_tab = np.concatenate(tabs).astype(np.uint32)
        b.payload = np.concatenate(pays).astype(np.uint8)
    return b


def region(x0, y0, x1, y1, mvx_q=0, mvy_q=0, ref=0, kind=abi.MB_INTER, dyn_base=0):
    r = np.zeros(1, abi.REGION)
    r["x0"], r["y0"], r["x1"], r["y1"] = x0, y0, x1, y1
    r["mvx"], r["mvy"], r["ref"], r["kind"], r["dyn_base"] = mvx_q, mvy_q, ref, kind, dyn_base
    return r


def dyn_region_payload(key, n_dyn):
    """Appendix B dyn-MB payload for one frame: (dyn_info hints, bit-offset table, payload bytes)."""
    j = np.arange(n_dyn, dtype=np.uint64)
    lens = (np.uint64(8) + mix(0xC0FFEE, key, j) % np.uint64(593)).astype(np.int64)
    tab = np.zeros(n_dyn + 1, np.int64)
    np.cumsum(lens, out=tab[1:])
    total_bits = int(tab[-1])
    n_words = (total_bits + 63) // 64
    words = mix(0xBEEF, key, np.arange(n_words, dtype=np.uint64))
    pay = words.astype(">u8").view(np.uint8).copy()          # MSB-first bytes
    zero = np.nonzero(mix(0xD15C, key, j) % np.uint64(20) == 0)[0]
    if os.environ.get("H264SCROLL_NO_ZERO_MBS"):              # experiments: how much the emulation-prevention slow path costs
        zero = zero[:0]
    for z in zero:                                            # all-zero MBs force 00 00 0x sequences
        b0, b1 = int(tab[z]), int(tab[z + 1])
        y0, y1 = (b0 + 7) // 8, b1 // 8
        if y0 <= y1:
            if b0 & 7:
                pay[b0 // 8] &= (0xFF00 >> (b0 & 7)) & 0xFF
            pay[y0:y1] = 0
            if b1 & 7:
                pay[b1 // 8] &= 0xFF >> (b1 & 7)
        else:                                                 # inside one byte
            keep = ((0xFF00 >> (b0 & 7)) & 0xFF) | (0xFF >> (b1 & 7))
            pay[b0 // 8] &= keep
    pay = pay[: (total_bits + 7) // 8]
    pad = (-len(pay)) % 4
    pay = np.concatenate([pay, np.zeros(pad, np.uint8)])      # 4-byte granules per frame
    info = np.zeros(n_dyn, abi.HINT)
    inter = mix(0x1B, key, j) % np.uint64(10) < 7
    info["kind"] = np.where(inter, abi.MB_DYN_INTER, abi.MB_DYN_INTRA)
    info["mvx"] = np.where(inter, (mix(0x1C, key, j) % np.uint64(129)).astype(np.int64) - 64, 0)
    info["mvy"] = np.where(inter, (mix(0x1D, key, j) % np.uint64(129)).astype(np.int64) - 64, 0)
    info["ref"] = 0
    info["dyn"] = np.arange(n_dyn)
    return info, tab.astype(np.uint32), pay


_SPEED_PERMS = [(8, 20, 36), (8, 36, 20), (20, 8, 36), (20, 36, 8), (36, 8, 20), (36, 20, 8)]


def config3_regions(t, speeds=(8, 20, 36), with_dyn=True):
    """Config 3 (1920x1088, 120x68 MBs): chrome + three carousels + vertical two-reference grid + dyn rect."""
    regs = [region(0, 0, 120, 68)]
    for k, (r0, r1) in enumerate(((8, 20), (24, 36), (40, 52))):
        mvx_px = -((speeds[k] * t) % 497)
        regs.append(region(8, r0, 120, r1, mvx_q=4 * mvx_px))
    off = int(legacy_schedule(t + 1, 4, 256)[t])
    split = 52 + (256 - off) // 16
    regs.append(region(8, 52, 120, split, mvy_q=4 * off, ref=0))
    regs.append(region(8, split, 120, 68, mvy_q=4 * (off - 256), ref=1))
    if with_dyn:
        regs.append(region(80, 6, 105, 31, kind=abi.REGION_DYN, dyn_base=0))
    return np.concatenate(regs)


def make_config5(n_streams, frames_per_stream, stream0=0, with_dyn=True, flags=0):
    """Config 5: independent 1080p (120x68 MB) streams, config-3 layout with per-stream phase / speeds."""
    b = Batch(120, 68)
    b.name = "config5: %d streams x %d frames, 1920x1088, config-3 layout%s" % (
        n_streams, frames_per_stream, " + 25x25 dyn splice" if with_dyn else "")
    headers, regs, infos, tabs, pays = [], [], [], [], []
    for s in range(stream0, stream0 + n_streams):
        phase = int(mix(5, s, 0) % np.uint64(256))
        speeds = _SPEED_PERMS[int(mix(5, s, 1) % np.uint64(6))]
        for t in range(frames_per_stream):
            tt = t + phase
            headers.append(p_header_bits(2 + t))
            regs.append(config3_regions(tt, speeds, with_dyn))
            if with_dyn:
                i, tb, p = dyn_region_payload(s * 1000000 + tt, 625)
                infos.append(i); tabs.append(tb); pays.append(p)
    # slot: 104 bits worst case per inline MB is far above what these hints produce; 16 B/MB + payload bound
    slot = ((b.n_mbs * 4 + (625 * 76 * 3 // 2 if with_dyn else 0) + 256) + 15) & ~15
    return _finish(b, headers, regs, 2, flags, slot, (infos, tabs, pays) if with_dyn else None)


def make_config3(n_frames, with_dyn=True, flags=0):
    b = make_config5(1, n_frames, stream0=0, with_dyn=with_dyn, flags=flags)
    return b


def make_config4(n_frames, flags=0):
    """Config 4: 3840x2160 (240x135), per-row varying MVs, ref = (row/16)&1."""
    b = Batch(240, 135)
    b.name = "config4: %d frames 3840x2160 per-row MVs" % n_frames
    headers, regs = [], []
    rows = np.arange(135, dtype=np.uint64)
    for t in range(n_frames):
        mvy = 4 * ((mix(4, rows, t) % np.uint64(993)).astype(np.int64) - 496)
        r = np.zeros(135, abi.REGION)
        r["x0"], r["x1"] = 0, 240
        r["y0"], r["y1"] = np.arange(135), np.arange(135) + 1
        r["mvy"] = mvy
        r["ref"] = (np.arange(135) // 16) & 1
        headers.append(p_header_bits(2 + t))
        regs.append(r)
    slot = ((b.n_mbs * 6 + 256) + 15) & ~15
    return _finish(b, headers, regs, 2, flags, slot)


def make_config2(n_frames, flags=0):
    """Config 2: 1280x720 legacy scroll (0->720->0, speed 4) with a 25x25 dyn rect at MB (28,10);
    num_refs stays 2 here (the waypoint state machine is exercised through the composer API instead)."""
    b = Batch(80, 45)
    b.name = "config2: %d frames 1280x720 scroll + 25x25 dyn splice" % n_frames
    offs = legacy_schedule(n_frames, 4, 720)
    headers, regs, infos, tabs, pays = [], [], [], [], []
    for t in range(n_frames):
        off = int(offs[t])
        split = (720 - off) // 16
        rr = [region(0, 0, 80, split, mvy_q=4 * off, ref=0),
              region(0, split, 80, 45, mvy_q=4 * (off - 720), ref=1),
              region(28, 10, 53, 35, kind=abi.REGION_DYN, dyn_base=0)]
        headers.append(p_header_bits(2 + t))
        regs.append(np.concatenate(rr))
        i, tb, p = dyn_region_payload(t, 625)
        infos.append(i); tabs.append(tb); pays.append(p)
    slot = ((b.n_mbs * 6 + 625 * 76 * 3 // 2 + 256) + 15) & ~15
    return _finish(b, headers, regs, 2, flags, slot, (infos, tabs, pays))


def make_random(mb_w, mb_h, n_frames, seed, num_refs=2, flags=0, p_dyn=0.1, mv_range=64, n_regions=12,
                zero_payload_pct=5, hdr_bits=None):
    """Adversarial random frames for parity tests: random occluding rectangles, random dyn rectangles,
    zero-heavy payloads.  Region based so that the rasteriser is exercised as well."""
    rng = np.random.default_rng(seed)
    b = Batch(mb_w, mb_h)
    b.name = "random %dx%d seed %d" % (mb_w, mb_h, seed)
    headers, regs, infos, tabs, pays = [], [], [], [], []
    for t in range(n_frames):
        rr = [region(0, 0, mb_w, mb_h, int(rng.integers(-mv_range, mv_range + 1)),
                     int(rng.integers(-mv_range, mv_range + 1)), int(rng.integers(0, num_refs)))]
        info_parts, n_dyn = [], 0
        for _ in range(n_regions):
            x0, y0 = int(rng.integers(0, mb_w)), int(rng.integers(0, mb_h))
            x1, y1 = int(rng.integers(x0 + 1, mb_w + 1)), int(rng.integers(y0 + 1, mb_h + 1))
            if rng.random() < p_dyn and (x1 - x0) * (y1 - y0) <= 900 and n_dyn + (x1 - x0) * (y1 - y0) < 60000:
                cnt = (x1 - x0) * (y1 - y0)
                rr.append(region(x0, y0, x1, y1, kind=abi.REGION_DYN, dyn_base=n_dyn))
                inf = np.zeros(cnt, abi.HINT)
                inter = rng.random(cnt) < 0.7
                inf["kind"] = np.where(inter, abi.MB_DYN_INTER, abi.MB_DYN_INTRA)
                inf["mvx"] = np.where(inter, rng.integers(-mv_range, mv_range + 1, cnt), 0)
                inf["mvy"] = np.where(inter, rng.integers(-mv_range, mv_range + 1, cnt), 0)
                inf["ref"] = np.where(inter, rng.integers(0, num_refs, cnt), 0)
                inf["dyn"] = np.arange(n_dyn, n_dyn + cnt)
                info_parts.append(inf)
                n_dyn += cnt
            else:
                big = rng.random() < 0.15
                lim = 30000 if big else mv_range
                rr.append(region(x0, y0, x1, y1, int(rng.integers(-lim, lim + 1)), int(rng.integers(-lim, lim + 1)),
                                 int(rng.integers(0, num_refs))))
        lens = rng.integers(0, 700, n_dyn)
        tab = np.zeros(n_dyn + 1, np.int64)
        np.cumsum(lens, out=tab[1:])
        nbytes = (int(tab[-1]) + 7) // 8
        pay = rng.integers(0, 256, nbytes + ((-nbytes) % 4), dtype=np.uint8)
        bits = np.unpackbits(pay)
        for z in np.nonzero(rng.integers(0, 100, n_dyn) < zero_payload_pct)[0]:
            bits[int(tab[z]):int(tab[z + 1])] = 0
        # low-entropy stretches provoke 00 00 0x with x in 1..3 as well
        for z in np.nonzero(rng.integers(0, 100, n_dyn) < zero_payload_pct)[0]:
            seg = bits[int(tab[z]):int(tab[z + 1])]
            seg[:] = (rng.random(seg.size) < 0.04).astype(np.uint8)
        pay = np.packbits(bits)
        headers.append(hdr_bits if hdr_bits is not None else p_header_bits(2 + t))
        regs.append(np.concatenate(rr))
        infos.append(np.concatenate(info_parts) if info_parts else np.zeros(0, abi.HINT))
        tabs.append(tab.astype(np.uint32))
        pays.append(pay)
    slot = ((b.n_mbs * 16 + sum(len(p) for p in pays) // max(n_frames, 1) * 2 + 4096) + 15) & ~15
    return _finish(b, headers, regs, num_refs, flags, slot, (infos, tabs, pays))


def make_config1(n_streams=1, n_frames=248):
    """Config 1 (experiments/scroll-encoder twin): 1280x720, offsets from 496 px at speed 1 between 0 and 704
    (experiments/scroll-encoder/src/main.c:387-415); that program writes the waypoint frame INSTEAD of the scroll frame
    (:418-424), num_refs goes 2 -> 3 behind it.  Slice headers and the two bands per frame come from the host library's
    state machine (include/h264_writer.h; checked against the reference CLI's stream in tests/test_gpu_legacy_api.py).
    n_streams copies of the stream make the batch larger than L2."""
    import ctypes as C
    from . import binding
    lib = binding.load_library()
    b = Batch(80, 45)
    b.name = "config1: %d x %d frames 1280x720 scroll-encoder twin (waypoint at 496 px, num_refs 2 -> 3)" % (n_streams, n_frames)
    cfg = binding.ComposerConfig()
    lib.composer_config_init(C.byref(cfg), 1280, 720)
    lib.composer_config_set_sps_params(C.byref(cfg), 4, 2, 4)
    lib.composer_config_set_pps_params(C.byref(cfg), 1, 1)
    cfg.frame_num = 2
    headers, regs, nrefs, nalh = [], [], [], []
    out5, hdr = (C.c_int * 5)(), (C.c_uint64 * 4)()
    for i in range(n_frames):
        pos = (i + 496) % (2 * 704)
        off = pos if pos < 704 else 2 * 704 - pos
        wf = int(lib.h264_needs_waypoint(C.byref(cfg), off))
        nbits = lib.h264_p_slice_header_bits(C.byref(cfg), wf, hdr)
        bits = "".join(format(int(hdr[w]), "064b") for w in range(4))[:nbits]
        lib.h264_scroll_regions(C.byref(cfg), off, wf, out5)
        split = max(0, min(int(out5[0]), 45))
        headers.append(bits)
        regs.append(np.concatenate([region(0, 0, 80, split, mvy_q=4 * int(out5[2]), ref=int(out5[1])),
                                    region(0, split, 80, 45, mvy_q=4 * int(out5[4]), ref=int(out5[3]))]))
        nrefs.append(2 + cfg.num_waypoints)
        nalh.append(0x41 if wf else 0x01)
        if wf:                                                   # register the waypoint (src/h264_writer.c:771-777)
            w = cfg.waypoints[cfg.num_waypoints]
            w.offset_px, w.long_term_idx, w.valid = off, 2 + cfg.num_waypoints, 1
            cfg.num_waypoints += 1
        cfg.frame_num += 1
    slot = ((b.n_mbs * 6 + 256) + 15) & ~15
    b = _finish(b, headers * n_streams, regs * n_streams, 2, 0, slot, None, np.array(nalh * n_streams, np.uint8))
    b.desc["num_refs"] = np.array(nrefs * n_streams, np.uint8)
    return b


def repeat_batch(b, times):
    """`times` copies of a region-list batch back to back (frame-indexed arrays re-based): the same frames as independent
    units of work, to make a batch larger than L2."""
    if times <= 1:
        return b
    n = b.n_frames
    r = Batch(b.mb_w, b.mb_h)
    r.name, r.slot_cap = b.name + " x%d" % times, b.slot_cap
    r.desc, r.rdesc = np.tile(b.desc, times), np.tile(b.rdesc, times)
    k = np.repeat(np.arange(times, dtype=np.uint64), n)
    r.desc["hint_off"] = np.arange(n * times, dtype=np.uint64) * np.uint64(b.hint_stride)
    r.desc["out_off"] = np.arange(n * times, dtype=np.uint64) * np.uint64(b.slot_cap)
    r.rdesc["hint_off"] = r.desc["hint_off"]
    r.rdesc["region_first"] = (r.rdesc["region_first"].astype(np.uint64) + k * np.uint64(len(b.regions))).astype(np.uint32)
    r.regions = np.tile(b.regions, times)
    if b.dyn is not None:
        r.dyn = np.tile(b.dyn, times)
        r.dyn["table_off"] = r.dyn["table_off"] + k * np.uint64(len(b.dyn_tab))
        r.dyn["payload_off"] = r.dyn["payload_off"] + k * np.uint64(len(b.payload))
        r.desc["dyn_id"] = (np.tile(b.desc["dyn_id"], times).astype(np.uint64) + k * np.uint64(n)).astype(np.uint32)
        r.rdesc["dyn_info_off"] = (r.rdesc["dyn_info_off"].astype(np.uint64) + k * np.uint64(len(b.dyn_info))).astype(np.uint32)
        r.dyn_info, r.dyn_tab, r.payload = np.tile(b.dyn_info, times), np.tile(b.dyn_tab, times), np.tile(b.payload, times)
    return r
