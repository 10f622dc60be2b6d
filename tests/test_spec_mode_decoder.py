"""
SURVEY 8(f)-3: the SPEC modes (spec-correct P_Skip + H.264 8.4.1.3 predictor) have no counterpart in the reference, so
nothing of the reference can pin them.  An independent decoder can: tests/p_slice_mv_decoder.py parses the slice data
(mb_skip_run, mb_type, ref_idx, mvd) and reconstructs every macroblock's (ref, mv) from the specification text alone.
The hints are the intended motion, so   decode(compose(hints, SKIP_SPEC | PRED_SPEC)) == hints   must hold for every
inter macroblock - skipped ones included.

The decoder itself is pinned first: it must reproduce the motion of the REFERENCE's own legacy scroll streams (two bands;
the reference's median3 / single-neighbour quirks do not fire there).
"""
import numpy as np
import pytest

import helpers
import p_slice_mv_decoder as dec
from h264_scroll_encoder_b200 import abi, workloads

SPEC = abi.SKIP_SPEC | abi.PRED_SPEC


def _split_nals(stream):
    b = stream.tobytes()
    idx, i = [], b.find(b"\x00\x00\x00\x01")
    while i >= 0:
        idx.append(i)
        i = b.find(b"\x00\x00\x00\x01", i + 4)
    return [b[s:(idx[k + 1] if k + 1 < len(idx) else len(b))] for k, s in enumerate(idx)]


def test_decoder_reads_the_references_legacy_stream(oracle):
    # 1280x720, offsets through the first waypoint: 2-ref and 3-ref frames, a waypoint frame with MMCO in its header
    offs = workloads.legacy_schedule(140, 4, 720)
    stream, sizes = oracle.legacy_stream(1280, 720, offs)          # == the reference's bytes (tests/test_oracle_pinned.py)
    nals = _split_nals(stream)
    assert len(nals) == len(sizes) == 141
    k, waypoints = 0, []
    for off in (int(o) for o in offs):
        for waypoint_frame in ([True, False] if (off and off % 496 == 0 and off not in waypoints) else [False]):
            nal = nals[k]; k += 1
            if k % 9 and not waypoint_frame and off not in (492, 496, 500):
                continue                                           # (a sample: the Python bit reader is slow)
            # header length: 31 bits with two refs; with a waypoint listed 5 + 1 more ue pair; the MMCO frame is longer
            nrefs = 2 + len(waypoints)
            hdr = 31 if nrefs == 2 else 31 + 6
            if waypoint_frame:
                hdr = 31 + (6 if waypoints else 0) + 1 + 5 + 5 + 5 + 1 - 1 + 1
            split = (720 - off) // 16
            top = (0, (0, 4 * off)) if not waypoints or off <= 496 else (2, (0, 4 * (off - 496)))
            bot = (1, (0, 4 * (off - 720)))
            got = None
            for h in range(hdr - 8, hdr + 9):                      # the header is the host's business: find its end
                try:
                    got = dec.decode_p_slice(nal, 80, 45, nrefs, h)
                    break
                except (AssertionError, ValueError):
                    continue
            assert got is not None, "frame at offset %d does not decode" % off
            for mb, (skipped, ref, mv) in enumerate(got):
                want = top if mb // 80 < split else bot
                assert not skipped and (ref, mv) == want, (off, mb, ref, mv, want)
            if waypoint_frame:
                waypoints.append(off)


def _opaque(b, hints, f):
    """Pre-encoded macroblocks of frame f as the decoder needs them: bit count, neighbour-visible (ref, mv)."""
    h, pay, tab = helpers.frame_inputs(b, hints, f)
    out = {}
    for mb in np.nonzero(h["kind"] != abi.MB_INTER)[0]:
        di = int(h["dyn"][mb])
        intra = h["kind"][mb] == abi.MB_DYN_INTRA
        out[int(mb)] = (int(tab[di + 1]) - int(tab[di]), -1 if intra else int(h["ref"][mb]),
                        (0, 0) if intra else (int(h["mvx"][mb]), int(h["mvy"][mb])))
    return h, out


def _check_frames(compose, b, po, frames):
    hints = helpers.dense_hints(po, b)
    n_skipped = 0
    for f in frames:
        nal = compose(b, hints, f)
        h, opaque = _opaque(b, hints, f)
        got = dec.decode_p_slice(nal, b.mb_w, b.mb_h, int(b.desc["num_refs"][f]), int(b.desc["hdr_nbits"][f]), opaque)
        for mb, (skipped, ref, mv) in enumerate(got):
            if mb in opaque:
                continue
            want = (int(h["ref"][mb]), (int(h["mvx"][mb]), int(h["mvy"][mb])))
            assert (ref, mv) == want, "%s frame %d MB %d: decoded %r, hinted %r (skipped %r)" % (b.name, f, mb, (ref, mv), want, skipped)
            n_skipped += skipped
    return n_skipped


CASES = [
    ("random 23x17, 3 refs, spliced rectangles", lambda: workloads.make_random(23, 17, 3, seed=5, num_refs=3, flags=SPEC, p_dyn=0.3, n_regions=14), [0, 1, 2]),
    ("random 64x40, 5 refs", lambda: workloads.make_random(64, 40, 2, seed=8, num_refs=5, flags=SPEC, p_dyn=0.0, n_regions=30), [0, 1]),
    ("random 9x300 (tall), 2 refs", lambda: workloads.make_random(9, 300, 1, seed=9, num_refs=2, flags=SPEC, p_dyn=0.2, n_regions=25), [0]),
    ("config 2 (720p + splice)", lambda: workloads.make_config2(40, flags=SPEC), [0, 13, 39]),
    ("config 3 (1088p mixed + splice)", lambda: workloads.make_config3(40, flags=SPEC), [0, 17]),
    ("config 4 (2160p per-row MVs)", lambda: workloads.make_config4(2, flags=SPEC), [1]),
    ("config 5 (another stream)", lambda: workloads.make_config5(1, 20, stream0=77, flags=SPEC), [19]),
]


@pytest.mark.parametrize("name,make,frames", CASES, ids=[c[0] for c in CASES])
def test_spec_mode_of_the_oracle_decodes_to_the_hints(po, oracle, name, make, frames):
    b = make()
    skipped = _check_frames(lambda b, hints, f: helpers.checker_frame(oracle, b, hints, f).tobytes(), b, po, frames)
    if "config" in name:
        assert skipped > 0                         # uniform regions over ref 0 must actually come out as P_Skip


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,frames", CASES[:5], ids=[c[0] for c in CASES[:5]])
def test_spec_mode_of_the_gpu_decodes_to_the_hints(gpu, po, name, make, frames):
    b = make()
    hints = helpers.dense_hints(po, b)
    out, _ = gpu.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
    _check_frames(lambda b, hints, f: out[f].tobytes(), b, po, frames)
