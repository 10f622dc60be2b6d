"""
Library contexts (include/composer_gpu.h, SURVEY 8e "one host thread + stream set per GPU"): a context per host thread,
on as many devices as the box has, composes the same bytes as the default context - concurrently, with nothing shared.
"""
import threading
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import binding, workloads

pytestmark = pytest.mark.gpu


def _stream_of(gc, b, n_chunks):
    hb, keep = gc.make_host_batch(b, pinned=False)
    cap = b.slot_cap * b.n_frames
    out, offs, sizes = np.zeros(cap, np.uint8), np.zeros(b.n_frames + 1, np.uint64), np.zeros(b.n_frames, np.uint32)
    gc.compose_regions_host(hb, out.ctypes.data, cap, offs.ctypes.data, sizes.ctypes.data, n_chunks)
    return out[:int(offs[-1])].copy(), offs


def test_contexts_on_threads_match_the_default_context(gpu, po, oracle):
    n_dev = max(1, int(gpu.lib.h264gpu_device_count()))
    batches = [workloads.make_random(60 + 7 * i, 33 + 3 * i, 6, seed=200 + i, num_refs=3, p_dyn=0.3, n_regions=16) for i in range(4)]
    want = [_stream_of(gpu, b, 2)[0] for b in batches]                     # default context, this thread
    hints = [helpers.dense_hints(po, b) for b in batches]
    for b, h, w in zip(batches, hints, want):                              # ... which matches the oracle
        ref = np.concatenate([helpers.checker_frame(oracle, b, h, f) for f in range(b.n_frames)])
        assert np.array_equal(w, ref)
    got, errs = [None] * len(batches), []

    def work(i):
        try:
            gc = binding.GpuComposer(i % n_dev, lib=gpu.lib, own_context=True)
            assert gpu.lib.h264gpu_ctx_device(gc.ctx) == i % n_dev
            for _ in range(3):                                             # several calls: the cached scratch is reused
                got[i] = _stream_of(gc, batches[i], 3)[0]
            gc.close()
        except Exception as e:                                             # noqa: BLE001
            errs.append((i, repr(e)))

    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for i in range(len(batches)):
        assert np.array_equal(got[i], want[i]), "context %d differs from the default context" % i
    # the default context still works from this thread afterwards
    assert np.array_equal(_stream_of(gpu, batches[0], 2)[0], want[0])


def test_a_thread_without_a_context_of_its_own_uses_the_default_device(gpu, po, oracle):
    # the current CUDA device is per host thread - a second thread must not land on device 0 by accident
    b = workloads.make_random(40, 24, 3, seed=77, num_refs=2, n_regions=8)
    want = _stream_of(gpu, b, 1)[0]
    out = []
    t = threading.Thread(target=lambda: out.append(_stream_of(gpu, b, 1)[0]))
    t.start(); t.join()
    assert len(out) == 1 and np.array_equal(out[0], want)
