"""
GPU parity tests proper: the CUDA path, called through the C ABI (include/composer_gpu.h), against the
oracle on the same seeded inputs.  Bit-exact: these are integer/byte computations.
"""
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def _compare(gpu, po, chk, b, hints=None):
    hints = helpers.dense_hints(po, b) if hints is None else hints
    frames, sizes = gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
    for f in range(b.n_frames):
        want = helpers.checker_frame(chk, b, hints, f)
        got = frames[f]
        if not np.array_equal(got, want):
            m = min(len(got), len(want))
            k = int(np.argmax(got[:m] != want[:m])) if m and (got[:m] != want[:m]).any() else m
            raise AssertionError("%s frame %d: GPU %d bytes vs oracle %d bytes, first difference at byte %d"
                                 % (b.name, f, len(got), len(want), k))


@pytest.mark.parametrize("flags", [0, abi.SKIP_SPEC, abi.PRED_SPEC, abi.SKIP_SPEC | abi.PRED_SPEC])
@pytest.mark.parametrize("seed", range(6))
def test_random_small_frames(gpu, po, oracle, seed, flags):
    nr = [1, 2, 3, 5, 10, 2][seed]
    b = workloads.make_random(5 + 9 * seed, 4 + 5 * seed, 3, seed=seed, num_refs=nr, flags=flags)
    _compare(gpu, po, oracle, b)


@pytest.mark.parametrize("flags", [0, abi.SKIP_SPEC | abi.PRED_SPEC])
def test_random_multi_tile_frames(gpu, po, oracle, flags):
    # > 1024 MBs per frame: tile boundaries, halo rows, carried partial words and EPB state
    b = workloads.make_random(120, 68, 3, seed=41, num_refs=3, flags=flags, n_regions=40)
    _compare(gpu, po, oracle, b)
    b = workloads.make_random(240, 135, 2, seed=42, num_refs=2, flags=flags, n_regions=60)
    _compare(gpu, po, oracle, b)


def test_against_reference_functions(gpu, po, ref):
    # the reference's own median3/get_mv_prediction/write_p16x16_mb/copy_bits behind the hinted loop
    for flags in (0, abi.SKIP_SPEC):
        b = workloads.make_random(80, 45, 3, seed=77, num_refs=3, flags=flags, n_regions=30)
        _compare(gpu, po, ref, b)


def test_edge_shapes(gpu, po, oracle):
    for (w, h) in ((1, 1), (1, 7), (7, 1), (2, 2), (33, 31), (1023, 2)):
        for flags in (0, 3):
            b = workloads.make_random(w, h, 2, seed=w * 100 + h, num_refs=2, flags=flags, n_regions=6)
            _compare(gpu, po, oracle, b)


def test_all_skipped_frame(gpu, po, oracle):
    # static frame: every MB is P_Skip in SPEC mode; the slice is header + one trailing mb_skip_run
    b = workloads.Batch(120, 68)
    regs = [workloads.region(0, 0, 120, 68)]
    b = workloads._finish(b, [workloads.p_header_bits(2)], regs, 2, abi.SKIP_SPEC | abi.PRED_SPEC, 4096)
    _compare(gpu, po, oracle, b)


def test_zero_heavy_payload_epb(gpu, po, oracle):
    b = workloads.make_random(60, 40, 4, seed=5, num_refs=2, flags=0, p_dyn=0.5, zero_payload_pct=60)
    _compare(gpu, po, oracle, b)


def test_config_generators(gpu, po, oracle):
    _compare(gpu, po, oracle, workloads.make_config2(6))
    _compare(gpu, po, oracle, workloads.make_config5(2, 3))
    _compare(gpu, po, oracle, workloads.make_config4(2))


def test_device_rasteriser_matches_checker(gpu, po):
    b = workloads.make_random(97, 53, 5, seed=9, num_refs=4, n_regions=50)
    want = helpers.dense_hints(po, b)
    d_rd, d_reg = gpu.to_device(b.rdesc), gpu.to_device(b.regions)
    d_info = gpu.to_device(b.dyn_info) if len(b.dyn_info) else None
    d_h = gpu.alloc(want.nbytes)
    gpu.lib.h264gpu_memset(d_h.ptr, 0, want.nbytes, None)
    gpu.raster_regions(d_rd, b.n_frames, b.n_mbs, d_reg, d_info, d_h)
    got = d_h.download(abi.HINT)
    for f in range(b.n_frames):
        s = slice(f * b.hint_stride, f * b.hint_stride + b.n_mbs)
        assert np.array_equal(got[s], want[s]), "rasteriser differs in frame %d" % f


def test_pack_frames_is_append_order(gpu, po, oracle):
    b = workloads.make_random(50, 30, 9, seed=3, num_refs=2, n_regions=20)
    hints = helpers.dense_hints(po, b)
    d_desc, d_h = gpu.to_device(b.desc), gpu.to_device(hints)
    d_dyn, d_tab, d_pay = gpu.to_device(b.dyn), gpu.to_device(b.dyn_tab), gpu.to_device(
        np.concatenate([b.payload, np.zeros(8, np.uint8)]))
    d_out = gpu.alloc(b.slot_cap * b.n_frames)
    d_packed = gpu.alloc(b.slot_cap * b.n_frames)
    d_sizes, d_offs = gpu.alloc(4 * b.n_frames), gpu.alloc(8 * (b.n_frames + 1))
    gpu.compose_batch(d_desc, b.n_frames, d_h, d_dyn, d_tab, d_pay, d_out, d_sizes)
    gpu.pack_frames(d_desc, d_sizes, b.n_frames, d_out, d_packed, d_offs)
    offs = d_offs.download(np.uint64)
    want = np.concatenate([helpers.checker_frame(oracle, b, hints, f) for f in range(b.n_frames)])
    assert int(offs[-1]) == len(want)
    got = d_packed.download(np.uint8, count=len(want))
    assert np.array_equal(got, want)


def test_overflowing_slot_is_reported(gpu, po, oracle):
    from h264_scroll_encoder_b200.binding import GpuError
    b = workloads.make_random(40, 30, 1, seed=1, num_refs=2, n_regions=10)
    hints = helpers.dense_hints(po, b)
    want = helpers.checker_frame(oracle, b, hints, 0)
    b.desc["out_cap"] = 64
    with pytest.raises(GpuError):
        gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload, out_bytes=4096)
    assert len(want) > 64


def test_foreign_dense_map_is_canonicalised(gpu, po, oracle):
    # a dense map written by somebody else: spliced intra MBs with arbitrary mv / ref 0 instead of mv 0 / ref 0xFF
    b = workloads.make_random(40, 24, 2, seed=11, num_refs=3, n_regions=12, p_dyn=0.5)
    hints = helpers.dense_hints(po, b)
    raw = hints.copy()
    intra = raw["kind"] == abi.MB_DYN_INTRA
    assert intra.any()
    raw["mvx"][intra] = 77
    raw["mvy"][intra] = -5
    raw["ref"][intra] = 0
    d_h = gpu.to_device(raw)
    gpu.canonicalize_hints(d_h, len(raw))
    gpu.sync()
    assert np.array_equal(d_h.download(abi.HINT), hints)
    _compare(gpu, po, oracle, b, hints=raw)          # compose_host canonicalises on the device; the oracle goes by `kind`


@pytest.mark.parametrize("w,h", [(1, 1), (1, 9), (2, 3), (9, 1), (17, 2), (263, 3), (300, 2), (1023, 1)])
def test_ragged_and_extreme_widths(gpu, po, oracle, w, h):
    # widths that are not multiples of the 8-MB segment, single rows/columns, frames wider than a full hint tile
    for flags in (0, abi.SKIP_SPEC | abi.PRED_SPEC):
        b = workloads.make_random(w, h, 2, seed=w * 31 + h, num_refs=3, flags=flags, n_regions=6)
        _compare(gpu, po, oracle, b)


def _all_spliced_batch(w, h, n_frames, flags=0):
    """Every macroblock of the frame spliced: several queue rounds per tile and tiles larger than the window."""
    b = workloads.Batch(w, h)
    b.name = "all spliced %dx%d" % (w, h)
    headers, regs, infos, tabs, pays = [], [], [], [], []
    for t in range(n_frames):
        headers.append(workloads.p_header_bits(2 + t))
        regs.append(workloads.region(0, 0, w, h, kind=abi.REGION_DYN, dyn_base=0))
        i, tb, p = workloads.dyn_region_payload(7000 + t, w * h)
        infos.append(i); tabs.append(tb); pays.append(p)
    slot = (w * h * 80 + 4096 + 15) & ~15
    return workloads._finish(b, headers, regs, 2, flags, slot, (infos, tabs, pays))


@pytest.mark.parametrize("w,h", [(64, 48), (120, 68)])
def test_all_spliced_frames_overflow_queue_and_window(gpu, po, oracle, w, h):
    for flags in (0, abi.SKIP_SPEC | abi.PRED_SPEC):
        _compare(gpu, po, oracle, _all_spliced_batch(w, h, 2, flags))


def test_config5_full_size_properties(gpu, po, oracle):
    """BASELINE config 5 at scale (32 streams x 64 frames = 2048 frames of 1920x1088) through the device-pointer
    calls: structural properties on every frame, byte parity with the oracle on a sample."""
    b = workloads.make_config5(32, 64)
    n = b.n_frames
    d_desc, d_rd, d_reg = gpu.to_device(b.desc), gpu.to_device(b.rdesc), gpu.to_device(b.regions)
    d_info, d_dyn, d_tab = gpu.to_device(b.dyn_info), gpu.to_device(b.dyn), gpu.to_device(b.dyn_tab)
    d_pay = gpu.to_device(np.concatenate([b.payload, np.zeros(16, np.uint8)]))
    d_h = gpu.alloc(8 * b.hint_stride * n)
    d_out, d_packed = gpu.alloc(b.slot_cap * n), gpu.alloc(b.slot_cap * n // 2)
    d_sizes, d_offs = gpu.alloc(4 * n), gpu.alloc(8 * (n + 1))
    outs = []
    for _ in range(2):                                    # idempotence: the same batch twice gives the same bytes
        gpu.raster_regions(d_rd, n, b.n_mbs, d_reg, d_info, d_h)
        gpu.compose_batch(d_desc, n, d_h, d_dyn, d_tab, d_pay, d_out, d_sizes)
        gpu.pack_frames(d_desc, d_sizes, n, d_out, d_packed, d_offs)
        offs = d_offs.download(np.uint64)
        outs.append((d_sizes.download(np.uint32), offs, d_packed.download(np.uint8, count=int(offs[-1]))))
    (sizes, offs, stream), (sizes2, offs2, stream2) = outs
    assert np.array_equal(sizes, sizes2) and np.array_equal(stream, stream2)
    assert (sizes <= b.slot_cap).all() and (sizes > 5).all()
    assert np.array_equal(offs[1:] - offs[:-1], sizes.astype(np.uint64))         # packed in append order, no gaps
    starts = offs[:-1].astype(np.int64)
    sc = np.stack([stream[starts + k] for k in range(5)], axis=1)
    assert (sc == np.array([0, 0, 0, 1, 0x01], np.uint8)).all()                  # start code + NAL header of a P slice
    # emulation prevention: behind the start code no 00 00 0x (x <= 2) anywhere inside a NAL unit
    z = stream == 0
    bad = np.nonzero(z[:-2] & z[1:-1] & (stream[2:] <= 2))[0]
    assert np.isin(bad, np.concatenate([starts, starts + 1])).all()
    assert (stream[(offs[1:] - 1).astype(np.int64)] != 0).all()                  # rbsp_stop_one_bit: no NAL ends in 00
    hints = d_h.download(abi.HINT)
    for f in np.random.default_rng(5).choice(n, 24, replace=False):
        want = helpers.checker_frame(oracle, b, hints, int(f))
        got = stream[int(offs[f]):int(offs[f + 1])]
        assert np.array_equal(got, want), "frame %d" % f


@pytest.mark.parametrize("w,h,seed", [(120, 68, 1), (120, 68, 2), (120, 68, 3), (240, 135, 4), (96, 200, 5), (64, 64, 6)])
def test_multi_tile_frames_that_outgrow_the_window(gpu, po, oracle, w, h, seed):
    """Several tiles per frame, queue items carried from tile to tile, and frames larger than one 32 KB window (the
    window is emitted in the middle of the frame, in front of a tile or in the middle of one)."""
    b = workloads.make_random(w, h, 2, seed=900 + seed, num_refs=2, p_dyn=0.5, n_regions=24, zero_payload_pct=5)
    hints = helpers.dense_hints(po, b)
    want = [helpers.checker_frame(oracle, b, hints, f) for f in range(b.n_frames)]
    assert max(len(x) for x in want) > 20000
    _compare(gpu, po, oracle, b, hints)


@pytest.mark.parametrize("flags", [0, abi.SKIP_SPEC | abi.PRED_SPEC])
def test_nothing_is_written_outside_a_frames_bytes(gpu, po, oracle, flags):
    """Canaries: the kernel stores whole 16-byte groups where it can; no byte of a slot behind the frame's last byte, and
    no byte of the arena outside the slots, may change (compute-sanitizer is not available on the GPU pool)."""
    b = workloads.make_random(77, 41, 6, seed=31 + flags, num_refs=3, flags=flags, p_dyn=0.3, n_regions=18, zero_payload_pct=20)
    hints = helpers.dense_hints(po, b)
    d_desc, d_h = gpu.to_device(b.desc), gpu.to_device(hints)
    d_dyn, d_tab = gpu.to_device(b.dyn), gpu.to_device(b.dyn_tab)
    d_pay = gpu.to_device(np.concatenate([b.payload, np.zeros(16, np.uint8)]))
    arena = b.slot_cap * b.n_frames + 4096
    d_out, d_sizes = gpu.alloc(arena), gpu.alloc(4 * b.n_frames)
    gpu.check(gpu.lib.h264gpu_memset(d_out.ptr, 0xAA, arena, None))
    gpu.compose_batch(d_desc, b.n_frames, d_h, d_dyn, d_tab, d_pay, d_out, d_sizes)
    gpu.sync()
    out, sizes = d_out.download(np.uint8), d_sizes.download(np.uint32)
    for f in range(b.n_frames):
        o, n = int(b.desc["out_off"][f]), int(sizes[f])
        want = helpers.checker_frame(oracle, b, hints, f)
        assert n == len(want) and np.array_equal(out[o:o + n], want)
        assert (out[o + n:o + b.slot_cap] == 0xAA).all(), "frame %d wrote behind its last byte" % f
    assert (out[b.slot_cap * b.n_frames:] == 0xAA).all()
