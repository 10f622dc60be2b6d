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
