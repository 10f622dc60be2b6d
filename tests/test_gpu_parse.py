"""
GPU Annex-B scan (nal_parser_next, src/nal_parser.c:28-65, over a whole stream) against this library's host C parser
and, where oracle/_ref is built, the reference's.
"""
import ctypes as C
import os
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, workloads

pytestmark = pytest.mark.gpu


class NALUnit(C.Structure):
    _fields_ = [("nal_ref_idc", C.c_int), ("nal_unit_type", C.c_int), ("data", C.c_void_p), ("size", C.c_size_t),
                ("rbsp_size", C.c_size_t)]


class NALParser(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def _cpu_units(lib, stream):
    p, u, out = NALParser(), NALUnit(), []
    lib.nal_parser_init(C.byref(p), C.c_void_p(stream.ctypes.data), C.c_size_t(len(stream)))
    while lib.nal_parser_next(C.byref(p), C.byref(u)):
        out.append((u.data - stream.ctypes.data, u.size, u.nal_ref_idc, u.nal_unit_type))
    return out


def _check(gpu, stream, max_nals=None):
    stream = np.ascontiguousarray(stream, np.uint8)
    libs = [gpu.lib]
    ref = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_composer.so")
    if os.path.exists(ref):
        libs.append(C.CDLL(ref))
    want = _cpu_units(libs[0], stream)
    for lib in libs[1:]:
        assert _cpu_units(lib, stream) == want
    units, count = gpu.parse_annexb(stream, max_nals or max(len(want), 1) + 3)
    got = [(int(u["offset"]), int(u["size"]), int(u["nal_ref_idc"]), int(u["nal_unit_type"])) for u in units]
    assert count == len(want)
    assert got == want[:len(got)]


def test_parse_hand_made_streams(gpu):
    sc3, sc4 = [0, 0, 1], [0, 0, 0, 1]
    _check(gpu, np.array(sc4 + [0x67, 1, 2, 3] + sc4 + [0x68, 9] + sc3 + [0x65, 7, 7, 0, 0] + sc4 + [0x41, 5], np.uint8))
    _check(gpu, np.array(sc3 + [0x01, 0xFF] * 20 + [0, 0, 0] + sc3 + [0x61, 3, 0, 0, 3, 1, 9], np.uint8))      # trailing zeros, an EPB inside
    _check(gpu, np.array([9, 9, 9] + sc4 + [0x65] + [0xAB] * 40000 + sc4 + [0x41, 1], np.uint8))              # garbage in front, several chunks
    _check(gpu, np.array(sc4 + [0x65, 1], np.uint8))


def test_parse_composed_stream(gpu, po):
    # the packed Annex-B stream of a config-5 batch, start codes at arbitrary alignment, thousands of chunks
    b = workloads.make_config5(4, 16)
    hints = helpers.dense_hints(po, b)
    frames, _ = gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
    stream = np.concatenate(frames)
    _check(gpu, stream)
    # fewer records than units: the count still tells, and every written record - the last one too - is complete
    for max_nals in (1, 10, b.n_frames - 1):
        units, count = gpu.parse_annexb(stream, max_nals)
        assert count == b.n_frames and len(units) == max_nals
        want = _cpu_units(gpu.lib, stream)[:max_nals]
        got = [(int(u["offset"]), int(u["size"]), int(u["nal_ref_idc"]), int(u["nal_unit_type"])) for u in units]
        assert got == want


def _ref_parser_lib(gpu):
    ref = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_composer.so")
    lib = C.CDLL(ref) if os.path.exists(ref) else gpu.lib          # (the reference's own parser where oracle/_ref is built)
    lib.ebsp_to_rbsp.restype = C.c_size_t
    lib.ebsp_to_rbsp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.rbsp_to_ebsp.restype = C.c_size_t
    lib.rbsp_to_ebsp.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    return lib


@pytest.mark.parametrize("which", ["config2", "config3", "config4", "config5 spec modes", "config1"])
def test_every_configuration_parses_cleanly_with_the_reference_parser(gpu, po, which):
    """north_star: the output "must parse cleanly with the reference's nal_parser" - every named configuration, not only
    config 5: nal_parser_next (src/nal_parser.c:14-65) finds one P-slice unit per frame exactly where the frames lie, and
    ebsp_to_rbsp (:67-88) followed by rbsp_to_ebsp (src/nal.c:24-50) gives the unit's bytes back."""
    b = {"config2": lambda: workloads.make_config2(12), "config3": lambda: workloads.make_config3(6),
         "config4": lambda: workloads.make_config4(2),
         "config5 spec modes": lambda: workloads.make_config5(2, 4, flags=abi.SKIP_SPEC | abi.PRED_SPEC),
         "config1": lambda: workloads.make_config1(1, 24)}[which]()
    hints = helpers.dense_hints(po, b)
    frames, _ = gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
    stream = np.ascontiguousarray(np.concatenate(frames), np.uint8)
    lib = _ref_parser_lib(gpu)
    units = _cpu_units(lib, stream)
    assert len(units) == b.n_frames
    starts = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
    for f, (off, size, ref_idc, typ) in enumerate(units):
        assert typ == 1 and off == starts[f] + 5 and size == len(frames[f]) - 5, "frame %d" % f      # 00 00 00 01 + header byte
        assert ref_idc == (int(b.desc["nal_header"][f]) >> 5) & 3
        ebsp = np.ascontiguousarray(stream[off:off + size])
        rbsp = np.zeros(size + 8, np.uint8)
        n = lib.ebsp_to_rbsp(rbsp.ctypes.data, ebsp.ctypes.data, size)
        back = np.zeros(2 * size + 8, np.uint8)
        m = lib.rbsp_to_ebsp(back.ctypes.data, len(back), rbsp.ctypes.data, n)
        assert m == size and np.array_equal(back[:m], ebsp), "frame %d: escape round trip" % f
        assert rbsp[n - 1] != 0                                                                         # rbsp_stop_one_bit in the last byte
