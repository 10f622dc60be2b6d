"""
The reference's session API (include/composer.h, src/main.c:95-139) end to end:

  composer_init -> composer_write_header -> composer_write_scroll_frame(s) -> composer_get_output

against the byte streams the reference CLI wrote (tests/golden/reference_cli.json, SURVEY 8c: 3 569 503 bytes,
sha256 e630f3a6...), with the atlas the reference's own scroll-encoder test mode produced (committed xz-compressed:
tests/golden/atlas_1280x720.h264.xz, sha256 recorded beside the stream's).  The header NALs (SPS, PPS, the two rewritten
I-frames) are host code and are checked without a GPU; the P-frames need the device.

composer_write_hinted_frames (hints + spliced dynamic rectangle on top of the scroll frames, with the waypoint state
machine running underneath) is checked against the reference's own functions behind the hinted loop
(oracle/_ref/libref_hinted.so) - slice headers included, which come from the reference's header writers there.
"""
import ctypes as C
import hashlib
import json
import lzma
import os
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, binding, workloads

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CLI = json.load(open(os.path.join(GOLD, "reference_cli.json")))["composer_1280x720_n250_s4"]


@pytest.fixture(scope="module")
def atlas(tmp_path_factory):
    data = lzma.decompress(open(os.path.join(GOLD, "atlas_1280x720.h264.xz"), "rb").read())
    assert hashlib.sha256(data).hexdigest() == CLI["atlas_sha256"]
    p = tmp_path_factory.mktemp("atlas") / "ref_720.h264"
    p.write_bytes(data)
    return str(p).encode()


def _session(lib, atlas):
    c = binding.Composer()
    assert lib.composer_init(C.byref(c), atlas, atlas) == 0
    assert lib.composer_get_width(C.byref(c)) == 1280 and lib.composer_get_height(C.byref(c)) == 720
    lib.composer_write_header(C.byref(c))
    return c


def _output(lib, c):
    lib.composer_get_output_size.restype = C.c_size_t
    lib.composer_get_output.restype = C.c_void_p
    n = lib.composer_get_output_size(C.byref(c))
    return np.ctypeslib.as_array((C.c_uint8 * n).from_address(lib.composer_get_output(C.byref(c)))).copy()


def test_header_nals_match_the_reference_cli(built, atlas):
    lib = binding.load_library()
    c = _session(lib, atlas)
    out = _output(lib, c)
    assert hashlib.sha256(out.tobytes()).hexdigest() == CLI["header_sha256"]
    assert out[4:14].tobytes().hex() == CLI["sps_hex"]
    assert c.cfg.frame_num == 2
    lib.composer_finish(C.byref(c))


def _nal_census(lib, stream):
    """(type, ref_idc, size) of every NAL unit, by nal_parser_next of `lib`."""
    class NALUnit(C.Structure):
        _fields_ = [("nal_ref_idc", C.c_int), ("nal_unit_type", C.c_int), ("data", C.c_void_p), ("size", C.c_size_t)]

    class NALParser(C.Structure):
        _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]
    p, u, out = NALParser(), NALUnit(), []
    lib.nal_parser_init(C.byref(p), C.c_void_p(stream.ctypes.data), C.c_size_t(stream.size))
    while lib.nal_parser_next(C.byref(p), C.byref(u)):
        out.append((u.nal_unit_type, u.nal_ref_idc, int(u.size)))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
def test_whole_session_matches_the_reference_cli(gpu, atlas, batched):
    lib = gpu.lib
    c = _session(lib, atlas)
    offs = workloads.legacy_schedule(250, 4, 720)                  # src/main.c:108-120
    if batched:
        o = np.ascontiguousarray(offs, np.int32)
        assert lib.composer_write_scroll_frames(C.byref(c), C.c_void_p(o.ctypes.data), len(o)) == 0
    else:
        for off in offs:
            lib.composer_write_scroll_frame(C.byref(c), int(off))
    out = _output(lib, c)
    assert len(out) == CLI["bytes"]
    assert hashlib.sha256(out.tobytes()).hexdigest() == CLI["sha256"]
    # ... and it parses cleanly with the reference's own nal_parser (and with this library's)
    libs = [lib]
    ref = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_composer.so")
    if os.path.exists(ref):
        libs.append(C.CDLL(ref))
    for L in libs:
        census = _nal_census(L, out)
        assert len(census) == CLI["nals"]
        assert [t for t, _, _ in census[:4]] == [7, 8, 5, 1]
        assert all(t == 1 for t, _, _ in census[4:])
        assert sum(1 for _, r, _ in census[4:] if r == 2) == 1     # one waypoint frame (offset 496)
    lib.composer_finish(C.byref(c))


# ------------------------------------------------------------------ hinted frames on top of the scroll frames

def _bands(height, offset, waypoints, waypoint_frame):
    """Test-side restatement of the band / reference choice of src/h264_writer.c:555-588 (and :693-729 for a waypoint
    frame): (split_row, ref_top, mv_top_px, ref_bot, mv_bot_px).  waypoints = registered offsets, in order."""
    top_ref, top_mv, bot_ref, bot_mv = 0, offset, 1, offset - height
    if offset > 496:
        best = 0
        for i, w in enumerate(waypoints):
            if w <= offset and w > best and offset - w <= 496:
                best, top_ref, top_mv = w, 2 + i, offset - w
    if not waypoint_frame and bot_mv < -496:
        for i, w in enumerate(waypoints):
            if w > offset and offset - w >= -496:
                bot_ref, bot_mv = 2 + i, offset - w
                break
    return (height - offset) // 16, top_ref, top_mv, bot_ref, bot_mv


def _ref_header(ref, width, height, frame_num, waypoints, waypoint_frame):
    """Slice-header bits from the reference's own writers (oracle/ref_shim: refhint_slice_header)."""
    cfg = binding.ComposerConfig()
    L = ref.lib
    L.composer_config_init(C.byref(cfg), width, height)
    L.composer_config_set_sps_params(C.byref(cfg), 4, 2, 4)
    L.composer_config_set_pps_params(C.byref(cfg), 1, 1)
    cfg.frame_num = frame_num
    cfg.num_waypoints = len(waypoints)
    for i, w in enumerate(waypoints):
        cfg.waypoints[i].offset_px, cfg.waypoints[i].long_term_idx, cfg.waypoints[i].valid = w, 2 + i, 1
    buf = np.zeros(64, np.uint8)
    nbits = L.refhint_slice_header(C.byref(cfg), int(waypoint_frame), C.c_void_p(buf.ctypes.data), C.c_size_t(buf.size))
    words = np.zeros(4, np.uint64)
    for w in range(4):
        words[w] = int.from_bytes(buf[8 * w:8 * w + 8].tobytes(), "big")
    return words, nbits


@pytest.mark.gpu
def test_hinted_frames_with_waypoints_and_splice_match_the_reference_functions(gpu, po, ref, atlas):
    """BASELINE config 2 as specified: 1280x720 legacy scroll through the waypoint at 496 px (num_refs 2 -> 3, waypoint
    headers) WITH the 25x25-MB splice and a motion rectangle in the same frames."""
    lib = gpu.lib
    W, H, mb_w, mb_h = 1280, 720, 80, 45
    offs = [int(o) for o in workloads.legacy_schedule(250, 4, 720)][100:160]          # crosses 496 at the 25th frame
    c = _session(lib, atlas)
    head = len(_output(lib, c))
    frames = (binding.ComposerFrameHints * len(offs))()
    keep = []
    per_frame = []
    for t, off in enumerate(offs):
        info, tab, pay = workloads.dyn_region_payload(9000 + t, 625)
        mbi = (binding.ComposerMbInfo * 625)()
        for j in range(625):
            mbi[j].mv_x_qpel, mbi[j].mv_y_qpel = int(info["mvx"][j]), int(info["mvy"][j])
            mbi[j].ref_idx, mbi[j].intra = int(info["ref"][j]), int(info["kind"][j] == abi.MB_DYN_INTRA)
        d = binding.ComposerDynRegion(28, 10, 53, 35, pay.ctypes.data, pay.size, tab.ctypes.data, C.cast(mbi, C.c_void_p))
        mr = (binding.ComposerMotionRegion * 2)()
        mr[0] = binding.ComposerMotionRegion(4, 2, 20, 9, -(3 * t) % 61 - 30, 0, 0)   # a carousel over RefA
        mr[1] = binding.ComposerMotionRegion(60, 30, 80, 45, 0, 5, 1)                 # a patch of RefB that occludes the bands
        keep += [info, tab, pay, mbi, d, mr]
        frames[t].offset_px, frames[t].regions, frames[t].n_regions = off, C.cast(mr, C.c_void_p), 2
        frames[t].dyn = C.cast(C.pointer(d), C.c_void_p)
        per_frame.append((info, tab, pay, mr))
    assert lib.composer_write_hinted_frames(C.byref(c), frames, len(offs)) == 0
    got = _output(lib, c)[head:]
    assert c.cfg.num_waypoints == 1 and c.cfg.frame_num == 2 + len(offs) + 1

    # ---- the same stream from the reference's functions: state machine restated here, arithmetic theirs
    waypoints, frame_num, want = [], 2, []

    def one(off, waypoint_frame, hints_on_top):
        nonlocal frame_num
        hdr, nbits = _ref_header(ref, W, H, frame_num, waypoints, waypoint_frame)
        split, rt, mt, rb, mb = _bands(H, off, waypoints, waypoint_frame)
        regs = [workloads.region(0, 0, mb_w, max(0, min(split, mb_h)), mvy_q=4 * mt, ref=rt),
                workloads.region(0, max(0, min(split, mb_h)), mb_w, mb_h, mvy_q=4 * mb, ref=rb)]
        info = tab = pay = None
        if hints_on_top is not None:
            info, tab, pay, mr = hints_on_top
            for r in mr:
                regs.append(workloads.region(r.mb_x0, r.mb_y0, r.mb_x1, r.mb_y1, mvx_q=4 * r.mv_x_px, mvy_q=4 * r.mv_y_px, ref=r.ref_idx))
            regs.append(workloads.region(28, 10, 53, 35, kind=abi.REGION_DYN, dyn_base=0))
        hints = po.raster_regions_numpy(np.concatenate(regs), mb_w, mb_h, info)
        want.append(ref.compose_frame(hints, mb_w, mb_h, 2 + len(waypoints), 0, hdr, nbits,
                                      0x41 if waypoint_frame else 0x01, pay, tab))
        frame_num += 1

    for t, off in enumerate(offs):
        if off != 0 and off % 496 == 0 and off not in waypoints:
            one(off, True, None)
            waypoints.append(off)
        one(off, False, per_frame[t])
    want = np.concatenate(want)
    assert len(got) == len(want) and np.array_equal(got, want)
    lib.composer_finish(C.byref(c))


def test_band_restatement_matches_the_host_library(built):
    """The test-side _bands() above against the product's h264_scroll_regions over the whole state space it is used in."""
    lib = binding.load_library()
    for wps in ([], [496], [496, 992]):
        cfg = binding.ComposerConfig()
        lib.composer_config_init(C.byref(cfg), 1280, 1600)
        cfg.num_waypoints = len(wps)
        for i, w in enumerate(wps):
            cfg.waypoints[i].offset_px, cfg.waypoints[i].long_term_idx, cfg.waypoints[i].valid = w, 2 + i, 1
        out = (C.c_int * 5)()
        for off in range(0, 1601, 7):
            for wf in (0, 1):
                lib.h264_scroll_regions(C.byref(cfg), off, wf, out)
                assert tuple(out) == _bands(1600, off, wps, bool(wf)), (wps, off, wf)
