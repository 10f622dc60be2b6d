"""
GPU tests of h264gpu_compose_regions_host: host region lists + payload in, packed Annex-B stream out,
pipelined over frame chunks.  Byte parity against the oracle for every chunking, pageable and pinned inputs.
"""
import ctypes as C
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, workloads
from h264_scroll_encoder_b200.binding import GpuError

pytestmark = pytest.mark.gpu


def _run(gpu, b, n_chunks, pinned, stream_cap=None, want_sizes=True):
    hb, keep = gpu.make_host_batch(b, pinned=pinned)
    cap = int(b.slot_cap) * b.n_frames if stream_cap is None else stream_cap
    out = np.full(cap + 64, 0xAA, np.uint8)
    offs = np.zeros(b.n_frames + 1, np.uint64)
    sizes = np.zeros(b.n_frames, np.uint32) if want_sizes else None
    gpu.compose_regions_host(hb, out.ctypes.data, cap, offs.ctypes.data, None if sizes is None else sizes.ctypes.data, n_chunks)
    for k in keep:
        if isinstance(k, tuple):
            gpu.lib.h264gpu_host_free(k[0])
    return out, offs, sizes


@pytest.mark.parametrize("n_chunks", [0, 1, 2, 3, 7, 64])
@pytest.mark.parametrize("pinned", [False, True])
def test_stream_matches_oracle(gpu, po, oracle, n_chunks, pinned):
    b = workloads.make_random(50, 30, 11, seed=21, num_refs=3, n_regions=16)
    hints = helpers.dense_hints(po, b)
    want = [helpers.checker_frame(oracle, b, hints, f) for f in range(b.n_frames)]
    out, offs, sizes = _run(gpu, b, n_chunks, pinned)
    assert int(offs[0]) == 0 and int(offs[-1]) == sum(len(w) for w in want)
    for f in range(b.n_frames):
        assert int(sizes[f]) == len(want[f])
        assert np.array_equal(out[int(offs[f]):int(offs[f + 1])], want[f]), "frame %d differs" % f
    assert (out[int(offs[-1]):] == 0xAA).all()                 # nothing written behind the stream


def test_spec_flags_and_no_splice(gpu, po, oracle):
    for flags, p_dyn in ((abi.SKIP_SPEC | abi.PRED_SPEC, 0.1), (0, 0.0)):
        b = workloads.make_random(33, 20, 5, seed=5 + flags, num_refs=2, flags=flags, p_dyn=p_dyn)
        hints = helpers.dense_hints(po, b)
        out, offs, _ = _run(gpu, b, 2, True, want_sizes=False)
        for f in range(b.n_frames):
            assert np.array_equal(out[int(offs[f]):int(offs[f + 1])], helpers.checker_frame(oracle, b, hints, f))


def test_config5_chunks_agree(gpu, po, oracle):
    b = workloads.make_config5(3, 8)
    hints = helpers.dense_hints(po, b)
    a = _run(gpu, b, 1, True)
    c = _run(gpu, b, 5, True)
    assert np.array_equal(a[1], c[1]) and np.array_equal(a[0][:int(a[1][-1])], c[0][:int(c[1][-1])])
    for f in (0, 7, 13, 23):
        assert np.array_equal(a[0][int(a[1][f]):int(a[1][f + 1])], helpers.checker_frame(oracle, b, hints, f))


def test_errors_are_reported(gpu, po):
    b = workloads.make_random(40, 30, 4, seed=1, num_refs=2, n_regions=10)
    with pytest.raises(GpuError):                              # stream buffer too small
        _run(gpu, b, 2, False, stream_cap=100)
    b2 = workloads.make_random(40, 30, 4, seed=1, num_refs=2, n_regions=10)
    b2.desc["out_cap"] = 64                                    # slots too small
    with pytest.raises(GpuError):
        _run(gpu, b2, 2, False)
    b3 = workloads.make_random(40, 30, 4, seed=1, num_refs=2, n_regions=10)
    b3.rdesc["region_first"][2] = 10 ** 6                      # descriptor out of range
    with pytest.raises(GpuError):
        _run(gpu, b3, 2, False)
