"""
pyoracle.py - TEST INFRASTRUCTURE: ctypes access to the CPU checkers.

  Oracle   oracle/libscroll_oracle.so       this repo's plain-C restatement ("port")
  Ref      oracle/_ref/libref_hinted.so     the reference's own functions compiled in place
                                            (+ the hint-driven outer loop of ref_shim/ref_hinted.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product library never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(verbose=False):
    """Build the checkers.  The _ref targets are (re)built only where /root/reference exists."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _p(a):
    return a.ctypes.data if a is not None else None


class _Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)


class Oracle(_Lib):
    """Plain-C restatement (oracle/scroll_oracle.c)."""
    prefix = "so_"

    def __init__(self, path=None):
        super().__init__(path or os.path.join(HERE, "libscroll_oracle.so"))
        L = self.lib
        L.so_compose_frame.restype = C.c_size_t
        L.so_compose_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.so_legacy_stream.restype = C.c_size_t
        L.so_legacy_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_int]
        L.so_frame_stages.restype = None
        L.so_frame_stages.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.so_rbsp_to_ebsp.restype = C.c_size_t
        L.so_rbsp_to_ebsp.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.so_time_hinted.restype = C.c_double
        L.so_time_hinted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]

    def compose_frame(self, hints, mb_w, mb_h, num_refs, flags, hdr, hdr_nbits, nal_header=0x01,
                      payload=None, dyn_tab=None, cap=None):
        hints = np.ascontiguousarray(hints)
        hdr = np.ascontiguousarray(hdr, dtype=np.uint64)
        cap = cap or (64 + mb_w * mb_h * 16 + (payload.size if payload is not None else 0) * 2 + 64)
        out = np.zeros(cap, dtype=np.uint8)
        n = self.lib.so_compose_frame(_p(hints), mb_w, mb_h, num_refs, flags, _p(hdr), hdr_nbits, nal_header,
                                      _p(payload), payload.size if payload is not None else 0,
                                      _p(dyn_tab), _p(out), out.size)
        return out[:n].copy()

    def legacy_stream(self, width, height, offsets, deblock=1, cap=64 << 20):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        out = np.zeros(cap, dtype=np.uint8)
        sizes = np.zeros(2 * len(offsets) + 8, dtype=np.uint32)
        n = C.c_int(0)
        tot = self.lib.so_legacy_stream(width, height, deblock, _p(offsets), len(offsets), _p(out), out.size,
                                        _p(sizes), C.byref(n), sizes.size)
        return out[:tot].copy(), sizes[:n.value].copy()

    def frame_stages(self, hints, mb_w, mb_h, num_refs, flags, dyn_tab=None):
        hints = np.ascontiguousarray(hints)
        n = mb_w * mb_h
        px = np.zeros(n, np.int16); py = np.zeros(n, np.int16); bl = np.zeros(n, np.uint32)
        self.lib.so_frame_stages(_p(hints), mb_w, mb_h, num_refs, flags, _p(dyn_tab), _p(px), _p(py), _p(bl))
        return px, py, bl

    def rbsp_to_ebsp(self, rbsp):
        rbsp = np.ascontiguousarray(rbsp, dtype=np.uint8)
        out = np.zeros(rbsp.size * 3 // 2 + 8, np.uint8)
        n = self.lib.so_rbsp_to_ebsp(_p(out), out.size, _p(rbsp), rbsp.size)
        return out[:n].copy()

    def time_hinted(self, hints, n_frames, mb_w, mb_h, num_refs, flags, hdr, hdr_nbits,
                    payload=None, dyn_tab=None, tab_stride=0, reps=1):
        nbytes = C.c_uint64(0)
        hdr = np.ascontiguousarray(hdr, dtype=np.uint64)
        fn = getattr(self.lib, self.prefix + "time_hinted")
        t = fn(_p(hints), n_frames, mb_w, mb_h, num_refs, flags, _p(hdr), hdr_nbits,
               _p(payload), payload.size if payload is not None else 0, _p(dyn_tab), tab_stride, reps,
               C.byref(nbytes))
        return t, nbytes.value


class Ref(_Lib):
    """The reference's own code (compiled in place) behind the hinted outer loop."""
    prefix = "refhint_"

    def __init__(self, path=None):
        super().__init__(path or os.path.join(HERE, "_ref", "libref_hinted.so"))
        L = self.lib
        L.refhint_compose_frame.restype = C.c_size_t
        L.refhint_compose_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                            C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_void_p, C.c_size_t]
        L.refhint_legacy_stream.restype = C.c_size_t
        L.refhint_legacy_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                            C.c_void_p, C.c_void_p, C.c_int]
        L.refhint_time_hinted.restype = C.c_double
        L.refhint_time_hinted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.refhint_time_legacy.restype = C.c_double
        L.refhint_time_legacy.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def compose_frame(self, hints, mb_w, mb_h, num_refs, flags, hdr, hdr_nbits, nal_header=0x01,
                      payload=None, dyn_tab=None, cap=None):
        hints = np.ascontiguousarray(hints)
        hdr = np.ascontiguousarray(hdr, dtype=np.uint64)
        cap = cap or (64 + mb_w * mb_h * 16 + (payload.size if payload is not None else 0) * 2 + 64)
        out = np.zeros(cap, dtype=np.uint8)
        scratch = np.zeros(cap, dtype=np.uint8)
        n = self.lib.refhint_compose_frame(_p(hints), mb_w, mb_h, num_refs, flags, _p(hdr), hdr_nbits, nal_header,
                                           _p(payload), payload.size if payload is not None else 0,
                                           _p(dyn_tab), _p(out), out.size, _p(scratch), scratch.size)
        return out[:n].copy()

    def legacy_stream(self, width, height, offsets, deblock=1, cap=64 << 20):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        out = np.zeros(cap, dtype=np.uint8)
        sizes = np.zeros(2 * len(offsets) + 8, dtype=np.uint32)
        n = C.c_int(0)
        tot = self.lib.refhint_legacy_stream(width, height, deblock, _p(offsets), len(offsets), _p(out), out.size,
                                             _p(sizes), C.byref(n), sizes.size)
        return out[:tot].copy(), sizes[:n.value].copy()

    time_hinted = Oracle.time_hinted

    def time_legacy(self, width, height, offsets, reps=1):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        nbytes = C.c_uint64(0); nals = C.c_uint64(0)
        t = self.lib.refhint_time_legacy(width, height, _p(offsets), len(offsets), reps,
                                         C.byref(nbytes), C.byref(nals))
        return t, nbytes.value, nals.value


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_hinted.so"))


def raster_regions_numpy(regions, mb_w, mb_h, dyn_info=None):
    """Checker for the device rasteriser: later regions occlude earlier ones; uncovered = (ref 0, mv 0);
    spliced intra macroblocks come out in canonical form."""
    from importlib import import_module
    abi = import_module("h264_scroll_encoder_b200.abi")
    hints = np.zeros((mb_h, mb_w), dtype=abi.HINT)
    for r in regions:
        x0, y0, x1, y1 = max(int(r["x0"]), 0), max(int(r["y0"]), 0), min(int(r["x1"]), mb_w), min(int(r["y1"]), mb_h)
        if x1 <= x0 or y1 <= y0:
            continue
        if r["kind"] == abi.MB_INTER:
            v = np.zeros((), dtype=abi.HINT)
            v["mvx"], v["mvy"], v["ref"], v["kind"], v["dyn"] = r["mvx"], r["mvy"], r["ref"], 0, 0
            hints[y0:y1, x0:x1] = v
        else:
            rw = int(r["x1"]) - int(r["x0"])
            for y in range(y0, y1):
                base = int(r["dyn_base"]) + (y - int(r["y0"])) * rw + (x0 - int(r["x0"]))
                hints[y, x0:x1] = dyn_info[base:base + (x1 - x0)]
    intra = hints["kind"] == abi.MB_DYN_INTRA        # canonical form (include/composer_gpu.h): mv 0, ref 0xFF
    hints["mvx"][intra] = 0
    hints["mvy"][intra] = 0
    hints["ref"][intra] = 0xFF
    return hints.reshape(-1)
