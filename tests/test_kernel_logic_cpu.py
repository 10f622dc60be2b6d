"""
The bit-level logic of the CUDA kernels, checked on the CPU: tests/emu/emu_compose.cpp walks the same tile
pipeline as csrc/compose_kernels.cu with the very same helpers (csrc/mb_codec.cuh compiled by g++), so entry
layout, word-owned gather, chunk carries and the closed-form emulation prevention are compared with the
oracle without a GPU - with tiny thread counts (8 macroblocks each) and window sizes that force every boundary case.
`tile` below is the number of emulated threads per tile, `ch` the window size in words.
"""
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, workloads


@pytest.fixture(scope="module")
def emu():
    return helpers.build_emu()


def _check(emu, po, oracle, b, tile, ch):
    hints = helpers.dense_hints(po, b)
    for f in range(b.n_frames):
        want = helpers.checker_frame(oracle, b, hints, f)
        got = helpers.emu_frame(emu, b, hints, f, tile, ch)
        assert np.array_equal(got, want), "%s frame %d tile %d ch %d" % (b.name, f, tile, ch)


@pytest.mark.parametrize("flags", [0, abi.SKIP_SPEC, abi.PRED_SPEC, abi.SKIP_SPEC | abi.PRED_SPEC])
@pytest.mark.parametrize("tile,ch", [(256, 6144), (16, 64), (4, 8), (2, 3), (3, 1)])
def test_emulated_pipeline_matches_oracle(emu, po, oracle, tile, ch, flags):
    for seed, (w, h, nr) in enumerate(((7, 5, 2), (33, 17, 3), (64, 9, 10), (1, 40, 1), (50, 1, 2))):
        b = workloads.make_random(w, h, 2, seed=200 + seed, num_refs=nr, flags=flags, n_regions=10,
                                  zero_payload_pct=30)
        _check(emu, po, oracle, b, tile, ch)


def test_emulated_pipeline_zero_heavy(emu, po, oracle):
    b = workloads.make_random(40, 24, 3, seed=5, num_refs=2, flags=0, p_dyn=0.5, zero_payload_pct=70)
    for tile, ch in ((256, 6144), (16, 16), (4, 4), (1, 5)):
        _check(emu, po, oracle, b, tile, ch)


def test_emulated_pipeline_config_generators(emu, po, oracle):
    _check(emu, po, oracle, workloads.make_config2(2), 256, 6144)
    _check(emu, po, oracle, workloads.make_config5(1, 2), 256, 6144)
    _check(emu, po, oracle, workloads.make_config5(1, 1), 256, 512)
    _check(emu, po, oracle, workloads.make_config4(1), 256, 6144)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (9, 1), (17, 2), (263, 3), (300, 2), (1023, 1)])
def test_emulated_pipeline_ragged_and_extreme_widths(emu, po, oracle, w, h):
    # widths that are not multiples of the 8-MB segment; frames wider than a full hint tile (fewer segments per tile)
    for flags in (0, abi.SKIP_SPEC | abi.PRED_SPEC):
        b = workloads.make_random(w, h, 2, seed=w * 31 + h, num_refs=3, flags=flags, n_regions=6)
        _check(emu, po, oracle, b, 256, 6144)
        _check(emu, po, oracle, b, 8, 16)
