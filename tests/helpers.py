"""Shared helpers for the test-suite (CPU emulation build, dense-hint expansion, oracle drivers)."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "emu_compose.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libemu.so")


def build_emu():
    deps = [EMU_SRC, os.path.join(ROOT, "h264-scroll-encoder_b200", "csrc", "mb_codec.cuh"),
            os.path.join(ROOT, "include", "composer_gpu.h")]
    if (not os.path.exists(EMU_LIB)) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-x", "c++",
                        "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "h264-scroll-encoder_b200", "csrc"),
                        EMU_SRC, "-o", EMU_LIB], check=True)
    lib = C.CDLL(EMU_LIB)
    lib.emu_compose_frame.restype = C.c_uint32
    lib.emu_compose_frame.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int]
    return lib


def dense_hints(po, b):
    """Expand a workloads.Batch's region lists with the numpy checker rasteriser."""
    from h264_scroll_encoder_b200 import abi
    hints = np.zeros(b.n_frames * b.hint_stride, abi.HINT)
    for f in range(b.n_frames):
        r0, rc = int(b.rdesc["region_first"][f]), int(b.rdesc["region_count"][f])
        d0 = int(b.rdesc["dyn_info_off"][f])
        di = b.dyn_info[d0:] if b.dyn_info is not None else None
        hints[f * b.hint_stride:f * b.hint_stride + b.n_mbs] = po.raster_regions_numpy(
            b.regions[r0:r0 + rc], b.mb_w, b.mb_h, di)
    return hints


def frame_inputs(b, hints, f):
    """(hints, payload, dyn_tab) views of frame f."""
    h = hints[f * b.hint_stride:f * b.hint_stride + b.n_mbs]
    if b.dyn is None:
        return h, None, None
    t0 = int(b.dyn["table_off"][f]); p0 = int(b.dyn["payload_off"][f])
    pay = b.payload[p0:p0 + int(b.dyn["payload_bytes"][f])]
    tab = b.dyn_tab[t0:t0 + int(b.dyn["n_mbs"][f]) + 1]
    return h, pay, tab


def checker_frame(chk, b, hints, f):
    """Compose frame f with an oracle-like object (pyoracle.Oracle or pyoracle.Ref)."""
    h, pay, tab = frame_inputs(b, hints, f)
    d = b.desc[f]
    return chk.compose_frame(h, b.mb_w, b.mb_h, int(d["num_refs"]), int(d["flags"]), d["hdr_bits"],
                             int(d["hdr_nbits"]), int(d["nal_header"]), pay, tab)


def emu_frame(emu, b, hints, f, tile, ch):
    out = np.zeros(int(b.desc["out_off"][f]) + b.slot_cap, np.uint8)
    p = lambda a: a.ctypes.data if a is not None else None
    desc = np.ascontiguousarray(b.desc[f:f + 1])
    n = emu.emu_compose_frame(p(desc), p(hints), p(b.dyn), p(b.dyn_tab), p(b.payload), p(out), tile, ch)
    o = int(b.desc["out_off"][f])
    return out[o:o + n].copy()
