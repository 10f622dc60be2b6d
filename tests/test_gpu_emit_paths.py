"""
GPU parity of the emit stage's two ways out (csrc/emit_stage.cuh): windows without insertions leave by the direct
128-bit stores, windows with insertions (or behind a byte count that is not a multiple of 16) by the scan / re-aligned
stores.  Frames of several windows whose emulation-prevention bytes all fall into the first window: the later windows
see every residue of "bytes inserted so far", 0 mod 16 included.  rbsp_to_ebsp: /root/reference/src/nal.c:24-50.
"""
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def _payload(key, n_dyn, n_zero):
    """dyn_region_payload with its all-zero macroblocks replaced by n_zero of the first n_dyn / 6 macroblocks."""
    info, tab, pay = workloads.dyn_region_payload(key, n_dyn)
    rng = np.random.default_rng(key)
    pay = rng.integers(1, 256, len(pay), dtype=np.uint8)           # no zero byte anywhere ...
    t = tab.astype(np.int64)
    for z in rng.choice(n_dyn // 6, n_zero, replace=False):        # ... but in these macroblocks
        b0, b1 = int(t[z]), int(t[z + 1])
        y0, y1 = (b0 + 7) // 8, b1 // 8
        if y0 < y1:
            pay[y0:y1] = 0
    return info, tab, pay


def _batch(w, h, zero_counts):
    b = workloads.Batch(w, h)
    b.name = "all spliced %dx%d, insertions in the first window only" % (w, h)
    headers, regs, infos, tabs, pays = [], [], [], [], []
    for t, nz in enumerate(zero_counts):
        headers.append(workloads.p_header_bits(2 + t))
        regs.append(workloads.region(0, 0, w, h, kind=abi.REGION_DYN, dyn_base=0))
        i, tb, p = _payload(9000 + t, w * h, nz)
        infos.append(i); tabs.append(tb); pays.append(p)
    slot = (w * h * 80 + 4096 + 15) & ~15
    return workloads._finish(b, headers, regs, 2, 0, slot, (infos, tabs, pays))


def test_windows_behind_insertions_take_both_ways_out(gpu, po, oracle):
    # 64x48 all-spliced macroblocks at ~300 bits each: four 32 KB windows per frame; 0..39 all-zero macroblocks up front
    b = _batch(64, 48, list(range(0, 40)))
    hints = helpers.dense_hints(po, b)
    frames, sizes = gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
    residues = set()
    for f in range(b.n_frames):
        want = helpers.checker_frame(oracle, b, hints, f)
        assert np.array_equal(frames[f], want), "frame %d" % f
        # bytes inserted in the frame's first 32 KB (what the second window starts behind)
        z = want == 0
        ins = np.nonzero(z[:-2] & z[1:-1] & (want[2:] == 3))[0]
        residues.add(int((ins < 32768).sum()) % 16)
    assert len(frames[0]) > 3 * 32768                              # several windows per frame
    assert 0 in residues and len(residues) > 4                      # the direct way and the re-aligned ways both ran


def _soup_without_splice(w, h, n_frames, seed, num_refs, flags, n_regions):
    """Random occluding motion rectangles over a background, no spliced region at all (no offset table for the frame)."""
    rng = np.random.default_rng(seed)
    b = workloads.Batch(w, h)
    b.name = "soup without splice %dx%d seed %d" % (w, h, seed)
    headers, regs = [], []
    for t in range(n_frames):
        rr = [workloads.region(0, 0, w, h, int(rng.integers(-64, 65)), int(rng.integers(-64, 65)), int(rng.integers(0, num_refs)))]
        for _ in range(n_regions):
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            x1, y1 = int(rng.integers(x0 + 1, w + 1)), int(rng.integers(y0 + 1, h + 1))
            lim = 30000 if rng.random() < 0.15 else 64
            rr.append(workloads.region(x0, y0, x1, y1, int(rng.integers(-lim, lim + 1)), int(rng.integers(-lim, lim + 1)),
                                       int(rng.integers(0, num_refs))))
        headers.append(workloads.p_header_bits(2 + t))
        regs.append(np.concatenate(rr))
    return workloads._finish(b, headers, regs, num_refs, flags, (w * h * 16 + 4096 + 15) & ~15)


@pytest.mark.parametrize("flags", [0, abi.SKIP_SPEC, abi.PRED_SPEC, abi.SKIP_SPEC | abi.PRED_SPEC])
def test_frames_without_a_spliced_region_in_every_mode(gpu, po, oracle, flags):
    """The whole-warp "rows of one motion" vote of the decide stage (frames without an offset table): per-row motion
    (config 4, where it holds in every warp), region soups without spliced macroblocks (where it holds in some warps and
    not in others), ragged widths (never), in all four skip / predictor modes.  Reference loop: src/h264_writer.c:595-646."""
    for b in (workloads.make_config4(2, flags=flags),
              _soup_without_splice(120, 68, 2, 301, 3, flags, 10),
              _soup_without_splice(80, 45, 3, 302, 2, flags, 2),
              _soup_without_splice(240, 30, 2, 304, 2, flags, 1),
              _soup_without_splice(83, 11, 2, 303, 2, flags, 4)):
        assert b.dyn is None
        hints = helpers.dense_hints(po, b)
        frames, sizes = gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
        for f in range(b.n_frames):
            assert np.array_equal(frames[f], helpers.checker_frame(oracle, b, hints, f)), "%s frame %d" % (b.name, f)


def test_misaligned_device_pointers_are_refused(gpu):
    """d_desc / d_hints / d_out travel as 128-bit words: h264gpu_compose_batch refuses pointers that are not 16-byte aligned
    instead of launching a kernel that would fault (include/composer_gpu.h)."""
    b = workloads.make_config4(1)
    d_desc, d_h = gpu.to_device(b.desc), gpu.alloc(8 * b.hint_stride + 64)
    d_out, d_sizes = gpu.alloc(b.slot_cap + 64), gpu.alloc(16)
    L = gpu.lib
    ok = L.h264gpu_compose_batch(d_desc.ptr, 0, d_h.ptr, None, None, None, d_out.ptr, d_sizes.ptr, None)
    assert ok == 0
    for args in ((d_desc.ptr, d_h.ptr, d_out.ptr + 4), (d_desc.ptr, d_h.ptr + 8, d_out.ptr), (d_desc.ptr + 4, d_h.ptr, d_out.ptr)):
        rc = L.h264gpu_compose_batch(args[0], 1, args[1], None, None, None, args[2], d_sizes.ptr, None)
        assert rc != 0 and b"16-byte aligned" in L.h264gpu_last_error()
    for x in (d_desc, d_h, d_out, d_sizes):
        x.free()
