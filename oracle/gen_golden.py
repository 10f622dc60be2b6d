"""
gen_golden.py - regenerates tests/golden/* from the UNMODIFIED reference compiled in place
(oracle/_ref, built by oracle/Makefile from /root/reference).  Run in the build container only:

    python oracle/gen_golden.py

Outputs (small, committed):
  tests/golden/legacy_streams.json   sizes + sha256 of P-frame streams written by the reference's
                                     h264_write_scroll_p_frame / h264_write_waypoint_p_frame in the
                                     composer's per-frame sequence, at 720p / 1088p / 2160p
  tests/golden/reference_cli.json    sha256 / size / NAL census of full streams written by the reference
                                     CLIs (composer, experiments/scroll-encoder), and known-answer NALs
  tests/golden/hinted_small.npz      seeded hinted frames (hints, payload, tables) with the bytes produced
                                     by the reference's own get_mv_prediction / write_p16x16_mb / copy_bits /
                                     nal_write_unit behind the hint-driven loop (ref_shim/ref_hinted.c)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from h264_scroll_encoder_b200 import workloads, abi  # noqa: E402
import oracle.pyoracle as po  # noqa: E402
import helpers  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(ROOT, "oracle", "_ref")

LEGACY_CASES = {
    "720p_s4": dict(width=1280, height=720, frames=250, speed=4),
    "1088p_s7": dict(width=1920, height=1088, frames=300, speed=7),
    "2160p_s16": dict(width=3840, height=2160, frames=120, speed=16),
    "720p_s1_waypoints": dict(width=1280, height=720, frames=800, speed=1),
}


def nal_census(data):
    """Split Annex-B on 4-byte start codes (every NAL the reference writes uses them)."""
    b = data.tobytes() if hasattr(data, "tobytes") else data
    idx, out, i = [], [], b.find(b"\x00\x00\x00\x01")
    while i >= 0:
        idx.append(i)
        i = b.find(b"\x00\x00\x00\x01", i + 4)
    for k, s in enumerate(idx):
        e = idx[k + 1] if k + 1 < len(idx) else len(b)
        out.append(b[s:e])
    return out


def main():
    po.build()
    assert po.have_ref(), "needs /root/reference"
    ref = po.Ref()
    os.makedirs(GOLD, exist_ok=True)

    legacy = {}
    for name, c in LEGACY_CASES.items():
        offs = workloads.legacy_schedule(c["frames"], c["speed"], c["height"])
        data, sizes = ref.legacy_stream(c["width"], c["height"], offs)
        legacy[name] = dict(c, bytes=int(len(data)), nals=int(len(sizes)),
                            waypoints=int(len(sizes) - c["frames"]),
                            sha256=hashlib.sha256(data.tobytes()).hexdigest(),
                            first_nal_hex=data[:24].tobytes().hex(),
                            nal_sizes_head=[int(s) for s in sizes[:6]])
    json.dump(legacy, open(os.path.join(GOLD, "legacy_streams.json"), "w"), indent=1)

    cli = {}
    with tempfile.TemporaryDirectory() as td:
        for (w, h) in ((1280, 720), (1920, 1088), (3840, 2160)):
            atlas = os.path.join(td, "ref_%d.h264" % h)
            subprocess.run([os.path.join(REFBIN, "h264_scroll_encoder"), "-t", "-w", str(w), "-H", str(h),
                            "-n", "1", "-o", atlas], check=True, stdout=subprocess.DEVNULL)
            outp = os.path.join(td, "comp_%d.h264" % h)
            n = 250 if h == 720 else 40
            subprocess.run([os.path.join(REFBIN, "composer"), "--ref-a", atlas, "--ref-b", atlas, "-n", str(n),
                            "-s", "4", "-o", outp], check=True, stdout=subprocess.DEVNULL)
            data = open(outp, "rb").read()
            nals = nal_census(data)
            cli["composer_%dx%d_n%d_s4" % (w, h, n)] = dict(
                bytes=len(data), nals=len(nals), sha256=hashlib.sha256(data).hexdigest(),
                sps_hex=nals[0][4:].hex(), pps_hex=nals[1][4:].hex(),
                first_p_nal_bytes=len(nals[4]) - 4, first_p_head_hex=nals[4][4:16].hex(),
                p_stream_sha256=hashlib.sha256(b"".join(nals[4:])).hexdigest(),
                header_sha256=hashlib.sha256(b"".join(nals[:4])).hexdigest(),
                atlas_sha256=hashlib.sha256(open(atlas, "rb").read()).hexdigest())
        outp = os.path.join(td, "cfg1.h264")
        subprocess.run([os.path.join(REFBIN, "h264_scroll_encoder"), "-t", "--width", "1280", "--height", "720",
                        "--frames", "248", "--color-a", "red", "--color-b", "blue", "-o", outp],
                       check=True, stdout=subprocess.DEVNULL)
        data = open(outp, "rb").read()
        nals = nal_census(data)
        cli["config1_scroll_encoder_1280x720_248"] = dict(
            bytes=len(data), nals=len(nals), sha256=hashlib.sha256(data).hexdigest(),
            p_stream_sha256=hashlib.sha256(b"".join(nals[4:])).hexdigest(),
            p_nal_headers=sorted(set(n[4] for n in nals[4:])))
    json.dump(cli, open(os.path.join(GOLD, "reference_cli.json"), "w"), indent=1)

    # hinted frames through the reference's own arithmetic (REFERENCE predictor; skip OFF and SPEC)
    arrays, meta = {}, []
    k = 0
    for seed, (w, h, nr) in enumerate(((9, 7, 2), (16, 12, 3), (23, 9, 1), (40, 30, 5), (31, 17, 10), (80, 45, 2))):
        for flags in (0, abi.SKIP_SPEC):
            b = workloads.make_random(w, h, 2, seed=1000 + seed, num_refs=nr, flags=flags,
                                      n_regions=10, zero_payload_pct=20)
            hints = helpers.dense_hints(po, b)
            for f in range(b.n_frames):
                hh, pay, tab = helpers.frame_inputs(b, hints, f)
                out = helpers.checker_frame(ref, b, hints, f)
                arrays["hints_%d" % k] = hh.view(np.uint8)
                arrays["payload_%d" % k] = pay
                arrays["tab_%d" % k] = tab
                arrays["hdr_%d" % k] = b.desc["hdr_bits"][f]
                arrays["out_%d" % k] = out
                meta.append(dict(k=k, mb_w=w, mb_h=h, num_refs=nr, flags=flags,
                                 hdr_nbits=int(b.desc["hdr_nbits"][f])))
                k += 1
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "hinted_small.npz"), **arrays)
    print("golden written:", os.listdir(GOLD))


if __name__ == "__main__":
    main()
