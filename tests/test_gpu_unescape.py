"""
GPU EBSP -> RBSP (ebsp_to_rbsp, src/nal_parser.c:67-88) against this library's host C implementation (round-trip
checked against the oracle by the CPU tests) and, where oracle/_ref is built, the reference's own function.
"""
import ctypes as C
import os
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _cpu(lib, e):
    out = np.zeros(len(e) + 8, np.uint8)
    lib.ebsp_to_rbsp.restype = C.c_size_t
    n = lib.ebsp_to_rbsp(C.c_void_p(out.ctypes.data), C.c_void_p(e.ctypes.data), C.c_size_t(len(e)))
    return out[:n]


def _libs(gpu):
    libs = [("host", gpu.lib)]
    ref = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_composer.so")
    if os.path.exists(ref):
        libs.append(("reference", C.CDLL(ref)))
    return libs


def _check(gpu, payloads):
    got = gpu.unescape_nals(payloads)
    for name, lib in _libs(gpu):
        for i, p in enumerate(payloads):
            want = _cpu(lib, p)
            assert len(got[i]) == len(want), "%s payload %d: %d bytes vs %d" % (name, i, len(got[i]), len(want))
            if not np.array_equal(got[i], want):
                raise AssertionError("%s payload %d differs at byte %d" % (name, i, int(np.argmax(got[i] != want))))


def test_unescape_small_and_edge_patterns(gpu):
    rng = np.random.default_rng(1)
    alphabet = np.array([0, 0, 0, 0, 1, 2, 3, 3, 3, 9], np.uint8)
    pay = [rng.choice(alphabet, size=n) for n in (1, 2, 3, 4, 5, 15, 16, 17, 18, 31, 32, 33, 100, 1000)]
    pay += [np.array(b, np.uint8) for b in ([0, 0, 3], [0, 0, 3, 0], [0, 0, 3, 4], [0, 0, 3, 3], [0, 0, 0, 3, 1], [3, 0, 0, 3, 0, 0, 3, 1],
                                            [0, 0, 3, 0, 0, 3, 0, 0, 3, 0, 0, 3, 0, 0, 3, 0, 0, 3, 2], [0] * 40, [0, 0, 3] * 20)]
    _check(gpu, pay)


def test_unescape_multi_chunk(gpu):
    rng = np.random.default_rng(2)
    alphabet = np.array([0, 0, 0, 1, 3, 3, 200], np.uint8)
    a = rng.choice(alphabet, size=150_001)
    b = rng.integers(1, 256, 140_000, dtype=np.uint8)
    for centre in (32768, 65536, 98304):                     # emulation prevention patterns right on the chunk borders
        for shift in range(-3, 4):
            b[centre + 16 * shift + shift:centre + 16 * shift + shift + 4] = (0, 0, 3, 1)
    c = np.tile(np.array([0, 0, 3], np.uint8), 40_000)       # a removal in every group, everywhere
    _check(gpu, [a, b, c, rng.choice(alphabet, size=32768), rng.choice(alphabet, size=32769), rng.choice(alphabet, size=65535)])


def test_unescape_round_trip_with_rewrite_sizes(gpu, oracle):
    rng = np.random.default_rng(3)
    rbsp = rng.choice(np.array([0, 0, 0, 1, 2, 3, 77], np.uint8), size=700_000)
    ebsp = oracle.rbsp_to_ebsp(rbsp)
    got = gpu.unescape_nals([ebsp])[0]
    assert np.array_equal(got, rbsp)                         # rbsp_to_ebsp followed by ebsp_to_rbsp is the identity
