"""
Pins the oracle (oracle/scroll_oracle.c, this repo's plain-C restatement) before anything trusts it:

  * against the golden fixtures under tests/golden/, which oracle/gen_golden.py produced from the UNMODIFIED
    reference compiled in place (oracle/_ref) - these run everywhere, /root/reference is not needed;
  * against the compiled reference itself (oracle/_ref) on fresh seeded inputs, where it is present.

CPU only.
"""
import hashlib
import json
import os
import numpy as np
import pytest

import helpers
from h264_scroll_encoder_b200 import abi, workloads

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hinted_golden():
    z = np.load(os.path.join(GOLD, "hinted_small.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.mark.parametrize("case", ["720p_s4", "1088p_s7", "2160p_s16", "720p_s1_waypoints"])
def test_legacy_stream_matches_reference_digest(oracle, case):
    # src/composer.c:255-264 sequence: optional waypoint frame, then the scroll frame, per offset
    g = json.load(open(os.path.join(GOLD, "legacy_streams.json")))[case]
    offs = workloads.legacy_schedule(g["frames"], g["speed"], g["height"])
    data, sizes = oracle.legacy_stream(g["width"], g["height"], offs)
    assert len(data) == g["bytes"]
    assert len(sizes) == g["nals"] and len(sizes) - g["frames"] == g["waypoints"]
    assert [int(s) for s in sizes[:6]] == g["nal_sizes_head"]
    assert data[:24].tobytes().hex() == g["first_nal_hex"]
    assert hashlib.sha256(data.tobytes()).hexdigest() == g["sha256"]


def test_known_answer_first_p_frame(oracle):
    # offset 0, frame_num 2, 2 refs: NAL 01 e5 57 68 95 ff ff ...; 2704 bytes + NAL header at 720p
    cli = json.load(open(os.path.join(GOLD, "reference_cli.json")))
    for (w, h) in ((1280, 720), (1920, 1088), (3840, 2160)):
        key = [k for k in cli if k.startswith("composer_%dx%d" % (w, h))][0]
        data, sizes = oracle.legacy_stream(w, h, np.zeros(1, np.int32))
        assert data[:4].tobytes() == b"\x00\x00\x00\x01"
        assert len(data) - 4 == cli[key]["first_p_nal_bytes"]
        assert data[4:16].tobytes().hex() == cli[key]["first_p_head_hex"]


def test_reference_cli_p_stream_digest(oracle):
    # the P-frames the reference CLI appended after its 4 header NALs (src/main.c:108-128 schedule, speed 4)
    cli = json.load(open(os.path.join(GOLD, "reference_cli.json")))
    for key, (w, h) in (("composer_1280x720_n250_s4", (1280, 720)), ("composer_1920x1088_n40_s4", (1920, 1088)),
                        ("composer_3840x2160_n40_s4", (3840, 2160))):
        n = int(key.split("_n")[1].split("_")[0])
        offs = workloads.legacy_schedule(n, 4, h)
        data, _ = oracle.legacy_stream(w, h, offs)
        assert hashlib.sha256(data.tobytes()).hexdigest() == cli[key]["p_stream_sha256"], key


def test_hinted_golden_frames(oracle):
    # reference's own get_mv_prediction / write_p16x16_mb / copy_bits / nal_write_unit behind the hinted loop
    z, meta = _hinted_golden()
    assert len(meta) == 24
    for m in meta:
        k = m["k"]
        hints = z["hints_%d" % k].view(abi.HINT)
        pay = z["payload_%d" % k]
        got = oracle.compose_frame(hints, m["mb_w"], m["mb_h"], m["num_refs"], m["flags"], z["hdr_%d" % k],
                                   m["hdr_nbits"], 0x01, pay if pay.size else None, z["tab_%d" % k])
        assert np.array_equal(got, z["out_%d" % k]), "golden hinted frame %d" % k


@pytest.mark.parametrize("flags", [0, abi.SKIP_SPEC])
@pytest.mark.parametrize("seed", range(5))
def test_oracle_equals_compiled_reference_on_random_hints(po, oracle, ref, seed, flags):
    nr = [1, 2, 3, 6, 10][seed]
    b = workloads.make_random(11 + 13 * seed, 6 + 7 * seed, 2, seed=500 + seed, num_refs=nr, flags=flags,
                              n_regions=14, zero_payload_pct=25)
    hints = helpers.dense_hints(po, b)
    for f in range(b.n_frames):
        assert np.array_equal(helpers.checker_frame(oracle, b, hints, f), helpers.checker_frame(ref, b, hints, f))


def test_oracle_equals_compiled_reference_on_config_generators(po, oracle, ref):
    for b in (workloads.make_config2(4), workloads.make_config3(3), workloads.make_config4(1),
              workloads.make_config5(2, 2)):
        hints = helpers.dense_hints(po, b)
        for f in range(b.n_frames):
            assert np.array_equal(helpers.checker_frame(oracle, b, hints, f),
                                  helpers.checker_frame(ref, b, hints, f)), b.name


def test_oracle_legacy_equals_compiled_reference(oracle, ref):
    for (w, h, n, s) in ((1280, 720, 60, 31), (640, 1600, 90, 37), (64, 48, 30, 5)):
        offs = workloads.legacy_schedule(n, s, h)
        a, sa = oracle.legacy_stream(w, h, offs)
        b, sb = ref.legacy_stream(w, h, offs)
        assert np.array_equal(sa, sb) and np.array_equal(a, b)


def test_rbsp_to_ebsp_closed_form_cases(oracle):
    # src/nal.c:24-50 automaton on adversarial zero-heavy strings; checked against a direct python automaton
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 7, 64, 1000):
        for _ in range(20):
            r = rng.choice(np.array([0, 0, 0, 1, 2, 3, 4, 255], np.uint8), size=n)
            out, zc = [], 0
            for v in r:
                if zc >= 2 and v <= 3:
                    out.append(3); zc = 0
                out.append(int(v))
                zc = zc + 1 if v == 0 else 0
            assert oracle.rbsp_to_ebsp(r).tolist() == out


def test_stage_outputs_reference_quirks(oracle):
    # median3(a,b,c) == min(max(a,b),c) (src/h264_writer.c:362-367): A=(20), B=(12), C=(-40) -> pred -40
    hints = np.zeros(6, abi.HINT)          # 3x2 frame; all ref 0
    hints["mvy"] = [12, -40, 0, 20, 7, 0]  # row 0: B=12 (x0), C=-40 (x1); row 1 x0=20 is A of (1,1)
    px, py, bl = oracle.frame_stages(hints, 3, 2, 2, 0)
    # MB (1,1): A = (0,1)=20, B = (1,0)=-40, C = (2,0)=0  -> min(max(20,-40),0) = 0
    assert py[4] == 0
    # MB (0,1): A unavailable, B=(0,0)=12, C=(1,0)=-40 -> two available, both match: min(max(0,12),-40) = -40
    assert py[3] == -40
