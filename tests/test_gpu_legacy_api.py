"""
The reference-shaped C API (include/h264_writer.h, composer.h) driving the GPU path: byte-identical with
the oracle's restatement of the composer sequence, and with the golden digests recorded from the compiled
reference (tests/golden/legacy_streams.json).
"""
import ctypes as C
import hashlib
import json
import os
import numpy as np
import pytest

from h264_scroll_encoder_b200 import binding, workloads

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "legacy_streams.json")


def _run_api(lib, width, height, offsets, batched):
    cfg = binding.ComposerConfig()
    lib.composer_config_init(C.byref(cfg), width, height)
    lib.composer_config_set_sps_params(C.byref(cfg), 4, 2, 4)
    lib.composer_config_set_pps_params(C.byref(cfg), 1, 1)
    cfg.frame_num = 2
    out = np.zeros(64 << 20, np.uint8)
    nw = binding.NALWriter()
    lib.nal_writer_init(C.byref(nw), C.c_void_p(out.ctypes.data), C.c_size_t(out.size), None, C.c_size_t(0))
    lib.h264_write_scroll_p_frame.restype = C.c_size_t
    lib.h264_write_waypoint_p_frame.restype = C.c_size_t
    sizes = []
    if batched:
        offs = np.ascontiguousarray(offsets, np.int32)
        made = C.c_int(0)
        k = lib.h264_write_scroll_sequence(C.byref(nw), C.byref(cfg), C.c_void_p(offs.ctypes.data), len(offs),
                                           C.byref(made))
        assert k == len(offs) + made.value
    else:
        for o in offsets:
            if lib.h264_needs_waypoint(C.byref(cfg), int(o)):
                sizes.append(lib.h264_write_waypoint_p_frame(C.byref(nw), C.byref(cfg), int(o)))
            sizes.append(lib.h264_write_scroll_p_frame(C.byref(nw), C.byref(cfg), int(o)))
        assert all(s > 0 for s in sizes)
    return out[:nw.output_pos].copy(), cfg


@pytest.mark.parametrize("case", ["720p_s4", "1088p_s7", "2160p_s16", "720p_s1_waypoints"])
def test_legacy_sequence_matches_golden_and_oracle(gpu, oracle, case):
    g = json.load(open(GOLDEN))[case]
    offs = workloads.legacy_schedule(g["frames"], g["speed"], g["height"])
    got, cfg = _run_api(gpu.lib, g["width"], g["height"], offs, batched=True)
    want, _ = oracle.legacy_stream(g["width"], g["height"], offs)
    assert np.array_equal(got, want)
    assert hashlib.sha256(got.tobytes()).hexdigest() == g["sha256"]
    assert len(got) == g["bytes"]
    assert cfg.num_waypoints == g["waypoints"]


def test_single_frame_calls_match_batched(gpu, oracle):
    offs = workloads.legacy_schedule(40, 31, 720)       # passes 496 exactly at i=16
    a, cfg_a = _run_api(gpu.lib, 1280, 720, offs, batched=False)
    b, cfg_b = _run_api(gpu.lib, 1280, 720, offs, batched=True)
    want, _ = oracle.legacy_stream(1280, 720, offs)
    assert np.array_equal(a, want) and np.array_equal(b, want)
    assert cfg_a.frame_num == cfg_b.frame_num and cfg_a.num_waypoints == cfg_b.num_waypoints


def test_config1_scroll_encoder_twin_p_frames(gpu):
    """BASELINE config 1 (experiments/scroll-encoder, 1280x720, 248 frames from offset 496, speed 1): that program emits
    the waypoint frame INSTEAD of the scroll frame (experiments/scroll-encoder/src/main.c:418-424).  Its P NAL units,
    recorded from the compiled reference CLI (tests/golden/reference_cli.json), through this library's frame writers."""
    gold = json.load(open(os.path.join(os.path.dirname(GOLDEN), "reference_cli.json")))["config1_scroll_encoder_1280x720_248"]
    lib = gpu.lib
    cfg = binding.ComposerConfig()
    lib.composer_config_init(C.byref(cfg), 1280, 720)
    lib.composer_config_set_sps_params(C.byref(cfg), 4, 2, 4)
    lib.composer_config_set_pps_params(C.byref(cfg), 1, 1)
    cfg.frame_num = 2
    out = np.zeros(16 << 20, np.uint8)
    nw = binding.NALWriter()
    lib.nal_writer_init(C.byref(nw), C.c_void_p(out.ctypes.data), C.c_size_t(out.size), None, C.c_size_t(0))
    lib.h264_write_scroll_p_frame.restype = C.c_size_t
    lib.h264_write_waypoint_p_frame.restype = C.c_size_t
    max_off = 720 - 16
    headers = set()
    for i in range(248):
        pos = (i + 496) % (2 * max_off)
        off = pos if pos < max_off else 2 * max_off - pos
        before = nw.output_pos
        if lib.h264_needs_waypoint(C.byref(cfg), off):
            assert lib.h264_write_waypoint_p_frame(C.byref(nw), C.byref(cfg), off) > 0
        else:
            assert lib.h264_write_scroll_p_frame(C.byref(nw), C.byref(cfg), off) > 0
        headers.add(int(out[before + 4]))
    got = out[:nw.output_pos]
    assert sorted(headers) == gold["p_nal_headers"]
    assert hashlib.sha256(got.tobytes()).hexdigest() == gold["p_stream_sha256"]
