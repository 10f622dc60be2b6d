"""GPU box: randomized parity soak.  Random frame shapes, region soups, flags, reference counts and zero-heavy payloads
through h264gpu_compose_host (dense hints) and h264gpu_compose_regions_host (region lists, chunked), byte for byte
against the oracle.  Usage: gpu_fuzz.py [n_cases] [seed0]."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.load_package()
import helpers
from h264_scroll_encoder_b200 import binding, workloads, abi


def run(n_cases, seed0, gc=None, po=None, orc=None, verbose=True):
    po = po or ge.load_oracle(); orc = orc or po.Oracle(); gc = gc or binding.GpuComposer(0)
    rng = np.random.default_rng(seed0)
    t0 = time.time(); bad = 0; frames = 0
    for c in range(n_cases):
        kind = rng.integers(0, 10)
        if kind < 6: w, h = int(rng.integers(1, 90)), int(rng.integers(1, 60))
        elif kind < 8: w, h = int(rng.integers(100, 300)), int(rng.integers(1, 12))     # wide: few segments per tile row
        elif kind < 9: w, h = int(rng.integers(1, 9)), int(rng.integers(60, 400))       # narrow and tall
        else: w, h = 120, 68
        nf = int(rng.integers(1, 5)); refs = int(rng.integers(1, 9)); flags = int(rng.integers(0, 4))
        p_dyn = float(rng.choice([0.0, 0.05, 0.2, 0.6])); zp = int(rng.choice([0, 5, 30, 100]))
        mvr = int(rng.choice([0, 3, 64, 2000])); nreg = int(rng.integers(0, 25))
        b = workloads.make_random(w, h, nf, seed=seed0 + c, num_refs=refs, flags=flags, p_dyn=p_dyn, mv_range=mvr,
                                  n_regions=nreg, zero_payload_pct=zp)
        hints = helpers.dense_hints(po, b)
        want = [helpers.checker_frame(orc, b, hints, f) for f in range(b.n_frames)]
        got, _ = gc.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
        hb, keep = gc.make_host_batch(b, pinned=False)
        cap = int(b.slot_cap) * b.n_frames
        out = np.zeros(cap + 64, np.uint8); offs = np.zeros(b.n_frames + 1, np.uint64)
        gc.compose_regions_host(hb, out.ctypes.data, cap, offs.ctypes.data, None, int(rng.integers(0, 5)))
        for f in range(b.n_frames):
            frames += 1
            ok1 = np.array_equal(got[f], want[f]); ok2 = np.array_equal(out[int(offs[f]):int(offs[f + 1])], want[f])
            if not (ok1 and ok2):
                bad += 1
                if verbose:
                    print("MISMATCH case %d seed %d frame %d: %dx%d refs %d flags %d p_dyn %.2f zero%% %d mv %d regions %d (dense %s, regions %s)"
                          % (c, seed0 + c, f, w, h, refs, flags, p_dyn, zp, mvr, nreg, ok1, ok2))
                    for name, g in (("dense", got[f]), ("regions", out[int(offs[f]):int(offs[f + 1])])):
                        g = np.asarray(g); n = min(len(g), len(want[f]))
                        d = np.nonzero(g[:n] != want[f][:n])[0]
                        print("   %s: %d bytes (want %d), %d differing bytes, first at %s, last at %s" %
                              (name, len(g), len(want[f]), len(d), d[0] if len(d) else None, d[-1] if len(d) else None))
    if verbose:
        print("fuzz: %d cases, %d frames, %d mismatches, %.1f s" % (n_cases, frames, bad, time.time() - t0))
    return bad, frames


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    sys.exit(1 if run(n, s0)[0] else 0)
