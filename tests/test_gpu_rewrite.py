"""
GPU slice rewrite (the reference's I-frame path, src/h264_writer.c:228-350: new slice header + copy_bits of the
macroblock data + nal_write_unit) against a numpy restatement of the bit string with the oracle's rbsp_to_ebsp
(pinned to the reference's), and against this library's host C implementation of h264_rewrite_idr_frame.
"""
import ctypes as C
import numpy as np
import pytest

from h264_scroll_encoder_b200 import abi, binding

pytestmark = pytest.mark.gpu


def _expected(oracle, hdr_bits, nal_header, src, start_bit, n_bits):
    bits = np.concatenate([np.array([int(c) for c in hdr_bits], np.uint8),
                           np.unpackbits(src)[start_bit:start_bit + n_bits]])
    rbsp = np.packbits(bits)                                   # zero padded to a byte, as bitwriter_get_size does
    return np.concatenate([np.array([0, 0, 0, 1, nal_header], np.uint8), oracle.rbsp_to_ebsp(rbsp)])


def _run(gpu, oracle, cases):
    desc = np.zeros(len(cases), abi.REWRITE_DESC)
    srcs, off, out_off = [], 0, 0
    for i, (hdr, nalh, src, start, n) in enumerate(cases):
        pad = (-len(src)) % 4
        srcs.append(np.concatenate([src, np.zeros(pad, np.uint8)]))
        cap = ((len(hdr) + n) // 8 * 3 // 2 + 64 + 15) & ~15
        desc[i]["hdr_bits"] = abi.hdr_words(hdr)
        desc[i]["src_off"], desc[i]["src_bytes"] = off, len(src)
        desc[i]["src_start_bit"], desc[i]["n_bits"] = start, n
        desc[i]["out_off"], desc[i]["out_cap"] = out_off, cap
        desc[i]["hdr_nbits"], desc[i]["nal_header"] = len(hdr), nalh
        off += len(src) + pad
        out_off += cap
    got, sizes = gpu.rewrite_slices(desc, np.concatenate(srcs), out_off)
    for i, (hdr, nalh, src, start, n) in enumerate(cases):
        want = _expected(oracle, hdr, nalh, src, start, n)
        assert int(sizes[i]) == len(want), "slice %d: %d bytes vs %d" % (i, sizes[i], len(want))
        if not np.array_equal(got[i], want):
            k = int(np.argmax(got[i][:len(want)] != want))
            raise AssertionError("slice %d differs at byte %d of %d" % (i, k, len(want)))


def _hdr(rng, n):
    return "".join(rng.choice(["0", "1"], n))


def test_rewrite_small_and_ragged(gpu, oracle):
    rng = np.random.default_rng(1)
    cases = []
    for nbytes in (1, 2, 5, 16, 17, 63, 64, 65, 1000):
        src = rng.integers(0, 256, nbytes, dtype=np.uint8)
        start = int(rng.integers(0, min(8 * nbytes, 40)))
        cases.append((_hdr(rng, int(rng.integers(1, 200))), 0x65, src, start, 8 * nbytes - start))
    cases.append((_hdr(rng, 37), 0x61, rng.integers(0, 256, 50, dtype=np.uint8), 13, 0))      # nothing to copy
    _run(gpu, oracle, cases)


def test_rewrite_multi_chunk_zero_heavy(gpu, oracle):
    rng = np.random.default_rng(2)
    cases = []
    # several 32 KB chunks; zero-heavy so that insertions shift every later chunk
    src = rng.choice(np.array([0, 0, 0, 0, 1, 2, 3, 200], np.uint8), size=150_000)
    cases.append((_hdr(rng, 29), 0x65, src, 21, 8 * len(src) - 21))
    # zero runs across chunk borders (even and odd lengths), all-zero input: an insertion every second byte
    z = np.zeros(100_000, np.uint8)
    cases.append((_hdr(rng, 31), 0x65, z, 3, 8 * len(z) - 3))
    z2 = rng.integers(1, 256, 140_000, dtype=np.uint8)
    for centre in (32768 - 9, 65536 - 8, 98304 - 7):
        z2[centre - 300:centre + 300 + (centre & 1)] = 0
    cases.append(("1" + _hdr(rng, 58), 0x61, z2, 0, 8 * len(z2)))
    _run(gpu, oracle, cases)


def test_rewrite_i_frame_sizes(gpu, oracle):
    rng = np.random.default_rng(3)
    # the reference's 1280x720 I_PCM atlas is 1.4 MB; bytes drawn so that 00 00 0x shows up regularly
    src = rng.integers(0, 256, 1_400_000, dtype=np.uint8)
    src[rng.integers(0, len(src), 20000)] = 0
    src[rng.integers(0, len(src) - 1, 5000)[:, None] + np.arange(2)] = 0
    _run(gpu, oracle, [(_hdr(rng, 26), 0x65, src, 19, 8 * len(src) - 19),
                       (_hdr(rng, 45), 0x61, src[:700_001], 30, 8 * 700_001 - 30)])


def test_rewrite_matches_host_idr_rewrite(gpu, oracle):
    """The same I slice through this library's host h264_rewrite_idr_frame (checked against the reference's streams by
    the CPU tests) and through the device path with the header bits the host function writes."""
    lib = gpu.lib
    rng = np.random.default_rng(4)
    cfg, pcfg = binding.ComposerConfig(), binding.ComposerConfig()
    lib.composer_config_init(C.byref(cfg), 64, 48)
    lib.composer_config_init(C.byref(pcfg), 64, 48)
    # a source IDR slice: first_mb ue(0)=1, slice_type ue(7)=0001000, pps ue(0)=1, frame_num u(4), idr_pic_id ue(0)=1,
    # [poc type 2: none], no_output 0, long_term 0, slice_qp_delta se(0)=1, [deblocking per cfg], then macroblock data
    hdr_src = "1" + "0001000" + "1" + "0000" + "1" + "0" + "0" + "1"
    if cfg.deblocking_filter_control_present_flag:
        hdr_src += "010"                                      # disable_deblocking_filter_idc = 1
    body = rng.integers(0, 256, 40000, dtype=np.uint8)
    body[rng.integers(0, len(body), 2000)] = 0
    bits = np.concatenate([np.array([int(c) for c in hdr_src], np.uint8), np.unpackbits(body)])
    rbsp = np.packbits(bits)
    out = np.zeros(2 * len(rbsp) + 64, np.uint8)
    tmp = np.zeros(2 * len(rbsp) + 64, np.uint8)
    nw = binding.NALWriter()
    lib.nal_writer_init(C.byref(nw), C.c_void_p(out.ctypes.data), C.c_size_t(out.size),
                        C.c_void_p(tmp.ctypes.data), C.c_size_t(tmp.size))
    lib.h264_rewrite_idr_frame.restype = C.c_size_t
    n = lib.h264_rewrite_idr_frame(C.byref(nw), C.byref(cfg), C.byref(pcfg), C.c_void_p(rbsp.ctypes.data), C.c_size_t(len(rbsp)))
    assert n > 0
    want = out[:n]
    # the header the host wrote = everything in front of the copied macroblock data
    start = len(hdr_src)
    n_bits = 8 * len(rbsp) - start
    new_hdr = "1" + "0001000" + "1" + "0" * cfg.log2_max_frame_num
    idr_id = cfg.idr_pic_id + 1
    m = idr_id.bit_length() - 1
    new_hdr += "0" * m + format(idr_id, "b")
    if cfg.pic_order_cnt_type == 0:
        new_hdr += "0" * cfg.log2_max_pic_order_cnt_lsb
    new_hdr += "01" + "1"
    if cfg.deblocking_filter_control_present_flag:
        new_hdr += "010"
    desc = np.zeros(1, abi.REWRITE_DESC)
    desc[0]["hdr_bits"] = abi.hdr_words(new_hdr)
    desc[0]["src_bytes"], desc[0]["src_start_bit"], desc[0]["n_bits"] = len(rbsp), start, n_bits
    desc[0]["out_cap"] = (2 * len(rbsp) + 64) & ~15
    desc[0]["hdr_nbits"], desc[0]["nal_header"] = len(new_hdr), 0x65
    got, sizes = gpu.rewrite_slices(desc, np.concatenate([rbsp, np.zeros((-len(rbsp)) % 4, np.uint8)]), int(desc[0]["out_cap"]))
    assert int(sizes[0]) == n and np.array_equal(got[0], want)
