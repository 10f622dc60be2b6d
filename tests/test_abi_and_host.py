"""
CPU-side checks of the drop-in boundary and the host logic that stays in C:

  * libh264scroll_b200.so loads and exports every function include/*.h declares (no compute calls);
  * without a GPU every hot-path entry point fails loudly (no CPU fallback exists);
  * the host pieces kept from the reference API - SPS/PPS, P-slice header bits, waypoint state machine,
    region derivation, bit writer, rbsp_to_ebsp, Annex-B parser - agree with the oracle and the reference's
    known answers.
"""
import ctypes as C
import json
import os
import re
import numpy as np
import pytest

from h264_scroll_encoder_b200 import abi, binding, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib(built):
    return binding.load_library()


def _declared_functions(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//.*", "", text)
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined", "sizeof")))


@pytest.mark.parametrize("header", sorted(f for f in os.listdir(INC) if f.endswith(".h")))
def test_every_declared_symbol_is_exported(lib, header):
    names = _declared_functions(os.path.join(INC, header))
    assert names, "no prototypes parsed from %s" % header
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "%s declares symbols the library does not export: %s" % (header, missing)


def test_header_pod_sizes_match_numpy_mirrors():
    text = open(os.path.join(INC, "composer_gpu.h")).read()
    src = '#include "composer_gpu.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(H264GpuMbHint),sizeof(H264GpuFrameDesc),sizeof(H264GpuDynDesc),sizeof(H264GpuRegion),' \
          'sizeof(H264GpuRasterDesc),sizeof(H264GpuHostBatch));return 0;}\n'
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I" + INC, os.path.join(td, "s.c"), "-o", os.path.join(td, "s")], check=True)
        out = subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [abi.HINT.itemsize, abi.FRAME_DESC.itemsize, abi.DYN_DESC.itemsize,
                                     abi.REGION.itemsize, abi.RASTER_DESC.itemsize, C.sizeof(binding.HostBatch)]
    assert "extern \"C\"" in text


def _no_gpu(lib):
    return lib.h264gpu_device_count() == 0


def test_no_cpu_fallback_without_a_device(lib):
    if not _no_gpu(lib):
        pytest.skip("a CUDA device is visible here")
    assert lib.h264gpu_init(0) == -1                     # H264GPU_ERR_NO_DEVICE
    assert lib.h264gpu_last_error()
    # device-pointer entry points refuse to run, they never compute on the host
    assert lib.h264gpu_compose_batch(None, 0, None, None, None, None, None, None, None) == -1
    desc = np.zeros(1, abi.FRAME_DESC)
    hints = np.zeros(4, abi.HINT)
    out = np.zeros(256, np.uint8)
    sizes = np.zeros(1, np.uint32)
    rc = lib.h264gpu_compose_host(desc.ctypes.data, 1, hints.ctypes.data, 4, None, 0, None, 0, None, 0,
                                  out.ctypes.data, out.size, sizes.ctypes.data)
    assert rc == -1
    hb = binding.HostBatch()                            # the pipelined host-batch call refuses as well
    hb.n_frames = 1
    offs = np.zeros(2, np.uint64)
    assert lib.h264gpu_compose_regions_host(C.byref(hb), out.ctypes.data, out.size, offs.ctypes.data, None, 0) == -1
    assert not out.any()
    with pytest.raises(binding.GpuError):
        binding.GpuComposer(0)
    # the reference-shaped frame writer reports failure (0 bytes) instead of producing bytes on the CPU
    cfg = binding.ComposerConfig()
    lib.composer_config_init(C.byref(cfg), 64, 48)
    buf = np.zeros(4096, np.uint8)
    nw = binding.NALWriter()
    lib.nal_writer_init(C.byref(nw), C.c_void_p(buf.ctypes.data), C.c_size_t(buf.size), None, C.c_size_t(0))
    lib.h264_write_scroll_p_frame.restype = C.c_size_t
    assert lib.h264_write_scroll_p_frame(C.byref(nw), C.byref(cfg), 0) == 0
    assert nw.output_pos == 0 and not buf.any()


# ------------------------------------------------------------------------------------------ host logic

def _cfg(lib, w, h, deblock=1):
    cfg = binding.ComposerConfig()
    lib.composer_config_init(C.byref(cfg), w, h)
    lib.composer_config_set_sps_params(C.byref(cfg), 4, 2, 4)
    lib.composer_config_set_pps_params(C.byref(cfg), 1, deblock)
    cfg.frame_num = 2
    return cfg


def test_sps_pps_known_answers(lib):
    cli = json.load(open(os.path.join(GOLD, "reference_cli.json")))
    lib.h264_generate_sps.restype = C.c_size_t
    lib.h264_generate_pps.restype = C.c_size_t
    for key, (w, h) in (("composer_1280x720_n250_s4", (1280, 720)), ("composer_1920x1088_n40_s4", (1920, 1088)),
                        ("composer_3840x2160_n40_s4", (3840, 2160))):
        buf = np.zeros(64, np.uint8)
        n = lib.h264_generate_sps(C.c_void_p(buf.ctypes.data), C.c_size_t(64), w, h)
        assert "67" + buf[:n].tobytes().hex() == cli[key]["sps_hex"]
        n = lib.h264_generate_pps(C.c_void_p(buf.ctypes.data), C.c_size_t(64))
        assert "68" + buf[:n].tobytes().hex() == cli[key]["pps_hex"]


def test_slice_header_and_waypoint_state_machine_match_oracle(lib, po, oracle):
    """Walk a long scroll through the host state machine and the oracle's; headers, regions and waypoints agree."""
    L = oracle.lib

    class SoStream(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("mb_w", C.c_int), ("mb_h", C.c_int),
                    ("log2_max_frame_num", C.c_int), ("poc_type", C.c_int), ("log2_max_poc_lsb", C.c_int),
                    ("deblock_present", C.c_int), ("frame_num", C.c_int), ("n_wp", C.c_int),
                    ("wp_offset", C.c_int * 8), ("wp_lt_idx", C.c_int * 8), ("wp_valid", C.c_int * 8)]

    for (w, h, speed, n) in ((1280, 720, 31, 120), (640, 4800, 62, 300), (64, 1600, 16, 250)):
        cfg = _cfg(lib, w, h)
        st = SoStream()
        L.so_stream_init(C.byref(st), w, h, 1)
        st.frame_num = 2
        offs = workloads.legacy_schedule(n, speed, h)
        mb_w, mb_h = w // 16, h // 16
        for o in offs:
            o = int(o)
            need = lib.h264_needs_waypoint(C.byref(cfg), o)
            assert need == L.so_needs_waypoint(C.byref(st), o)
            for wp in ([1, 0] if need else [0]):
                ha, hb = (C.c_uint64 * 4)(), (C.c_uint64 * 4)()
                na = lib.h264_p_slice_header_bits(C.byref(cfg), wp, ha)
                nb = L.so_slice_header(C.byref(st), wp, hb)
                assert na == nb and list(ha) == list(hb)
                r = (C.c_int * 5)()
                lib.h264_scroll_regions(C.byref(cfg), o, wp, r)
                hints = np.zeros(mb_w * mb_h, abi.HINT)
                L.so_legacy_hints(C.byref(st), o, wp, C.c_void_p(hints.ctypes.data))
                hh = hints.reshape(mb_h, mb_w)
                split = min(max(r[0], 0), mb_h)
                if split > 0:
                    assert (hh[:split]["ref"] == r[1]).all() and (hh[:split]["mvy"] == 4 * r[2]).all()
                if split < mb_h:
                    assert (hh[split:]["ref"] == r[3]).all() and (hh[split:]["mvy"] == 4 * r[4]).all()
                assert (hh["mvx"] == 0).all() and (hh["kind"] == 0).all()
                if wp:      # what write_one() does after a waypoint frame
                    L.so_register_waypoint(C.byref(st), o)
                    k = cfg.num_waypoints
                    if k < 8:
                        cfg.waypoints[k].offset_px, cfg.waypoints[k].long_term_idx, cfg.waypoints[k].valid = o, 2 + k, 1
                        cfg.num_waypoints = k + 1
                cfg.frame_num += 1
                st.frame_num += 1
        assert cfg.num_waypoints == st.n_wp


def test_two_ref_header_known_answer(lib):
    cfg = _cfg(lib, 1280, 720)
    h = (C.c_uint64 * 4)()
    n = lib.h264_p_slice_header_bits(C.byref(cfg), 0, h)
    assert n == 31
    bits = format(h[0], "064b")[:31]
    assert bits == workloads.p_header_bits(2)
    # NAL 01 | e5 57 68 95 ...: header bits followed by six '1' bits per MB
    assert int((bits + "1").ljust(32, "1"), 2) == 0xE5576895


def test_bitwriter_expgolomb_and_trailing_bits(lib):
    class BW(C.Structure):
        _fields_ = [("buffer", C.c_void_p), ("capacity", C.c_size_t), ("byte_pos", C.c_size_t),
                    ("bit_pos", C.c_int), ("current_byte", C.c_uint8)]
    buf = np.zeros(64, np.uint8)
    bw = BW()
    lib.bitwriter_init(C.byref(bw), C.c_void_p(buf.ctypes.data), C.c_size_t(64))
    want = ""
    for v in (0, 1, 2, 3, 6, 7, 254, 255, 65535):
        lib.bitwriter_write_ue(C.byref(bw), v)
        m = (v + 1).bit_length() - 1
        want += "0" * m + format(v + 1, "b")
    for v in (0, 1, -1, 2, -2, 1984, -1984):
        lib.bitwriter_write_se(C.byref(bw), v)
        k = 2 * v - 1 if v > 0 else -2 * v
        m = (k + 1).bit_length() - 1
        want += "0" * m + format(k + 1, "b")
    lib.bitwriter_write_trailing_bits(C.byref(bw))
    want += "1"
    want += "0" * (-len(want) % 8)
    lib.bitwriter_get_size.restype = C.c_size_t
    n = lib.bitwriter_get_size(C.byref(bw))
    assert n == len(want) // 8
    assert "".join(format(b, "08b") for b in buf[:n]) == want


def test_host_rbsp_to_ebsp_and_parser_round_trip(lib, oracle):
    lib.rbsp_to_ebsp.restype = C.c_size_t
    lib.ebsp_to_rbsp.restype = C.c_size_t
    rng = np.random.default_rng(11)
    for n in (0, 1, 5, 300, 4096):
        r = rng.choice(np.array([0, 0, 0, 1, 2, 3, 9], np.uint8), size=n)
        e = np.zeros(2 * n + 8, np.uint8)
        k = lib.rbsp_to_ebsp(C.c_void_p(e.ctypes.data), C.c_size_t(e.size), C.c_void_p(r.ctypes.data), C.c_size_t(n))
        assert np.array_equal(e[:k], oracle.rbsp_to_ebsp(r))
        back = np.zeros(n + 8, np.uint8)
        m = lib.ebsp_to_rbsp(C.c_void_p(back.ctypes.data), C.c_void_p(e.ctypes.data), C.c_size_t(k))
        assert m == n and np.array_equal(back[:m], r)


def test_reference_parser_accepts_oracle_stream(lib, oracle):
    """Structural acceptance: the kept nal_parser walks a P-frame stream: count, types, ref_idc."""
    class NALUnit(C.Structure):
        _fields_ = [("nal_ref_idc", C.c_int), ("nal_unit_type", C.c_int), ("data", C.c_void_p), ("size", C.c_size_t),
                    ("rbsp_size", C.c_size_t)]

    class NALParser(C.Structure):
        _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]
    offs = workloads.legacy_schedule(40, 31, 720)
    data, sizes = oracle.legacy_stream(1280, 720, offs)
    p = NALParser()
    lib.nal_parser_init(C.byref(p), C.c_void_p(data.ctypes.data), C.c_size_t(data.size))
    got = []
    u = NALUnit()
    while lib.nal_parser_next(C.byref(p), C.byref(u)):
        got.append((u.nal_ref_idc, u.nal_unit_type, u.size))
    assert len(got) == len(sizes)
    assert all(t == 1 for _, t, _ in got)
    assert sum(1 for r, _, _ in got if r == 2) == len(sizes) - len(offs)     # waypoint frames are references
    assert [s + 5 for _, _, s in got] == [int(s) for s in sizes]
