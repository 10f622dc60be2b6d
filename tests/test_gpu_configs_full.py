"""
BASELINE.json's configurations at their full frame counts on the GPU, byte-for-byte against the checkers:

  config 2   1280x720 legacy scroll 0 -> 720 -> 0 at speed 4, 250 frames, 25x25-MB splice in every frame
  config 3   1920x1088 chrome + carousels + two-reference grid + 25x25-MB splice, frames t = 0..255
  config 4   3840x2160, per-row varying MVs (every MB carries a non-zero mvd), frames t = 0..63

Every frame is compared with oracle/scroll_oracle.c (the port) and - where oracle/_ref was built - with the
reference's own functions behind the hinted loop (oracle/_ref/libref_hinted.so).  Integer/byte work: bit-exact.
"""
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import workloads

pytestmark = pytest.mark.gpu


def _full_compare(gpu, po, checkers, b, chunk=64):
    """Compose the whole batch through the host-buffer C-ABI call in chunks of frames; every frame against every checker."""
    hints = helpers.dense_hints(po, b)
    bad = []
    for f0 in range(0, b.n_frames, chunk):
        f1 = min(b.n_frames, f0 + chunk)
        desc = b.desc[f0:f1].copy()
        desc["out_off"] -= desc["out_off"][0]
        frames, sizes = gpu.compose_host(desc, hints, b.dyn, b.dyn_tab, b.payload)
        for f in range(f0, f1):
            got = frames[f - f0]
            for name, chk in checkers:
                want = helpers.checker_frame(chk, b, hints, f)
                if not np.array_equal(got, want):
                    bad.append((f, name, len(got), len(want)))
    assert not bad, "%s: %d frame/checker pairs differ, first %s" % (b.name, len(bad), bad[:4])


def _checkers(po, oracle):
    out = [("port", oracle)]
    if po.have_ref():
        out.append(("reference functions", po.Ref()))
    return out


def test_config3_all_256_frames(gpu, po, oracle):
    _full_compare(gpu, po, _checkers(po, oracle), workloads.make_config3(256))


def test_config4_all_64_frames(gpu, po, oracle):
    _full_compare(gpu, po, _checkers(po, oracle), workloads.make_config4(64), chunk=16)


def test_config2_all_250_frames(gpu, po, oracle):
    _full_compare(gpu, po, _checkers(po, oracle), workloads.make_config2(250))


def test_config5_two_streams_all_frames(gpu, po, oracle):
    # config 5 is config 3's layout per stream with a stream-specific phase and speed permutation
    _full_compare(gpu, po, _checkers(po, oracle), workloads.make_config5(2, 64, stream0=1000))
