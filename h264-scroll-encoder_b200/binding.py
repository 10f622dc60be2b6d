"""
ctypes binding of libh264scroll_b200.so.  No torch, no CPU fallback: if the library is missing or no
sm_100 device is visible, GpuComposer() raises.
"""
import ctypes as C
import os
import numpy as np
from . import abi

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libh264scroll_b200.so")

# every symbol include/*.h declares; tests check the built library exports all of them
GPU_SYMBOLS = [
    "h264gpu_init", "h264gpu_shutdown", "h264gpu_device_count", "h264gpu_last_error", "h264gpu_sm_count",
    "h264gpu_ctx_create", "h264gpu_ctx_destroy", "h264gpu_ctx_make_current", "h264gpu_ctx_device",
    "h264gpu_malloc", "h264gpu_free", "h264gpu_host_alloc", "h264gpu_host_free", "h264gpu_memcpy_h2d",
    "h264gpu_memcpy_d2h", "h264gpu_memset", "h264gpu_stream_create", "h264gpu_stream_destroy",
    "h264gpu_stream_sync", "h264gpu_event_create", "h264gpu_event_destroy", "h264gpu_event_record",
    "h264gpu_event_elapsed_ms", "h264gpu_compose_batch", "h264gpu_compose_batch_ex", "h264gpu_raster_regions", "h264gpu_pack_frames",
    "h264gpu_compose_host", "h264gpu_compose_regions_host", "h264gpu_launch_count", "h264gpu_canonicalize_hints",
    "h264gpu_rewrite_slices", "h264gpu_rewrite_chunk_bytes", "h264gpu_unescape_nals", "h264gpu_unescape_chunk_bytes",
    "h264gpu_parse_annexb", "h264gpu_parse_annexb_scratch_words"]
REFERENCE_API_SYMBOLS = [
    # include/composer.h
    "composer_init", "composer_get_width", "composer_get_height", "composer_write_header",
    "composer_write_scroll_frame", "composer_get_output_size", "composer_get_output",
    "composer_write_to_file", "composer_finish", "composer_write_scroll_frames", "composer_write_hinted_frames",
    # include/h264_writer.h
    "composer_config_init", "composer_config_set_sps_params", "composer_config_set_pps_params",
    "h264_generate_sps", "h264_generate_pps", "h264_rewrite_idr_frame", "h264_rewrite_as_non_idr_i_frame",
    "h264_write_scroll_p_frame", "h264_needs_waypoint", "h264_write_waypoint_p_frame",
    "h264_p_slice_header_bits", "h264_scroll_regions", "h264_write_scroll_sequence",
    # include/bitwriter.h
    "bitwriter_init", "bitwriter_write_bits", "bitwriter_write_bit", "bitwriter_write_ue", "bitwriter_write_se",
    "bitwriter_write_trailing_bits", "bitwriter_flush", "bitwriter_get_size", "bitwriter_get_bit_position",
    "bitwriter_is_byte_aligned", "bitreader_init", "bitreader_read_bits", "bitreader_read_bit",
    "bitreader_read_ue", "bitreader_read_se", "bitreader_get_bit_position", "bitreader_is_byte_aligned",
    "bitreader_get_remaining_bytes", "bitreader_get_pointer",
    # include/nal.h, include/nal_parser.h
    "nal_writer_init", "nal_write_unit", "nal_writer_get_size", "nal_writer_get_output", "rbsp_to_ebsp",
    "nal_parser_init", "nal_parser_next", "ebsp_to_rbsp", "parse_sps", "parse_pps"]


class GpuError(RuntimeError):
    pass


class HostBatch(C.Structure):
    """H264GpuHostBatch (include/composer_gpu.h)."""
    _fields_ = [("desc", C.c_void_p), ("n_frames", C.c_int), ("rdesc", C.c_void_p),
                ("regions", C.c_void_p), ("n_regions", C.c_size_t),
                ("dyn_info", C.c_void_p), ("n_dyn_info", C.c_size_t),
                ("dyn", C.c_void_p), ("n_dyn", C.c_int),
                ("dyn_tab", C.c_void_p), ("n_tab", C.c_size_t),
                ("payload", C.c_void_p), ("payload_bytes", C.c_size_t),
                ("n_hints", C.c_size_t), ("slots_bytes", C.c_size_t), ("max_mbs_per_frame", C.c_int)]


def load_library(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise GpuError("%s is missing - run __graft_entry__.build() (nvcc, sm_100a); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t
    L.h264gpu_last_error.restype = C.c_char_p
    L.h264gpu_ctx_create.restype = vp; L.h264gpu_ctx_create.argtypes = [ci]
    L.h264gpu_ctx_destroy.argtypes = [vp]; L.h264gpu_ctx_destroy.restype = None
    L.h264gpu_ctx_make_current.argtypes = [vp]
    L.h264gpu_ctx_device.argtypes = [vp]
    L.h264gpu_malloc.restype = vp; L.h264gpu_malloc.argtypes = [sz]
    L.h264gpu_free.argtypes = [vp]; L.h264gpu_free.restype = None
    L.h264gpu_host_alloc.restype = vp; L.h264gpu_host_alloc.argtypes = [sz]
    L.h264gpu_host_free.argtypes = [vp]; L.h264gpu_host_free.restype = None
    for f in ("h264gpu_memcpy_h2d", "h264gpu_memcpy_d2h"):
        getattr(L, f).argtypes = [vp, vp, sz, vp]
    L.h264gpu_memset.argtypes = [vp, ci, sz, vp]
    L.h264gpu_stream_create.restype = vp
    L.h264gpu_stream_destroy.argtypes = [vp]; L.h264gpu_stream_destroy.restype = None
    L.h264gpu_stream_sync.argtypes = [vp]
    L.h264gpu_event_create.restype = vp
    L.h264gpu_event_destroy.argtypes = [vp]; L.h264gpu_event_destroy.restype = None
    L.h264gpu_event_record.argtypes = [vp, vp]
    L.h264gpu_event_elapsed_ms.argtypes = [vp, vp]; L.h264gpu_event_elapsed_ms.restype = C.c_float
    L.h264gpu_compose_batch.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp]
    L.h264gpu_compose_batch_ex.argtypes = [vp, ci, vp, sz, vp, vp, vp, vp, vp, vp]
    L.h264gpu_raster_regions.argtypes = [vp, ci, ci, vp, vp, vp, vp]
    L.h264gpu_pack_frames.argtypes = [vp, vp, ci, vp, vp, vp, vp]
    L.h264gpu_canonicalize_hints.argtypes = [vp, sz, vp]
    L.h264gpu_rewrite_slices.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
    L.h264gpu_unescape_nals.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
    L.h264gpu_parse_annexb.argtypes = [vp, sz, vp, C.c_uint32, vp, vp]
    L.h264gpu_parse_annexb_scratch_words.argtypes = [sz, C.c_uint32]; L.h264gpu_parse_annexb_scratch_words.restype = sz
    L.h264gpu_compose_host.argtypes = [vp, ci, vp, sz, vp, ci, vp, sz, vp, sz, vp, sz, vp]
    L.h264gpu_compose_regions_host.argtypes = [C.POINTER(HostBatch), vp, sz, vp, vp, ci]
    L.h264gpu_launch_count.restype = C.c_uint64
    return L


def _ptr(a):
    return a.ctypes.data if a is not None else None


class DeviceBuffer:
    def __init__(self, gc, nbytes):
        self.gc, self.nbytes = gc, int(nbytes)
        self.ptr = gc.lib.h264gpu_malloc(max(self.nbytes, 16))
        if not self.ptr:
            raise GpuError("device malloc of %d bytes failed: %s" % (nbytes, gc.err()))

    def upload(self, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.gc.check(self.gc.lib.h264gpu_memcpy_h2d(self.ptr, _ptr(arr), arr.nbytes, stream))
        return self

    def download(self, dtype=np.uint8, count=None, offset=0, stream=None, sync=True):
        dtype = np.dtype(dtype)
        count = (self.nbytes - offset) // dtype.itemsize if count is None else count
        out = np.empty(count, dtype=dtype)
        self.gc.check(self.gc.lib.h264gpu_memcpy_d2h(_ptr(out), self.ptr + offset, out.nbytes, stream))
        if sync:
            self.gc.sync(stream)
        return out

    def free(self):
        if self.ptr:
            self.gc.lib.h264gpu_free(self.ptr)
            self.ptr = None


class GpuComposer:
    """Thin object over the C ABI (include/composer_gpu.h)."""

    def __init__(self, device=0, lib=None, own_context=False):
        """own_context=False: the library's default context (h264gpu_init).  own_context=True: a context of its own
        (h264gpu_ctx_create), made current for the calling thread - every thread that uses the object calls
        make_current() first."""
        self.lib = lib or load_library()
        self.ctx = None
        if own_context:
            self.ctx = self.lib.h264gpu_ctx_create(device)
            if not self.ctx:
                raise GpuError("h264gpu_ctx_create(%d) failed: %s" % (device, self.err()))
            self.make_current()
        else:
            rc = self.lib.h264gpu_init(device)
            if rc != 0:
                raise GpuError("h264gpu_init(%d) failed (%d): %s" % (device, rc, self.err()))
        self.device = device

    def make_current(self):
        self.check(self.lib.h264gpu_ctx_make_current(self.ctx))

    def close(self):
        if self.ctx:
            self.lib.h264gpu_ctx_destroy(self.ctx)
            self.ctx = None

    def err(self):
        return (self.lib.h264gpu_last_error() or b"").decode()

    def check(self, rc):
        if rc != 0:
            raise GpuError("libh264scroll_b200 call failed (%d): %s" % (rc, self.err()))

    def sync(self, stream=None):
        self.check(self.lib.h264gpu_stream_sync(stream))

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        return self.alloc(arr.nbytes).upload(arr)

    # ---- host-buffer call (what tests use for parity) ------------------------------------------
    def compose_host(self, desc, hints, dyn=None, dyn_tab=None, payload=None, out_bytes=None):
        desc = np.ascontiguousarray(desc, dtype=abi.FRAME_DESC)
        hints = np.ascontiguousarray(hints, dtype=abi.HINT)
        n = len(desc)
        if out_bytes is None:
            out_bytes = int((desc["out_off"] + desc["out_cap"]).max()) if n else 0
        out = np.zeros(out_bytes, np.uint8)
        sizes = np.zeros(n, np.uint32)
        dyn = None if dyn is None else np.ascontiguousarray(dyn, dtype=abi.DYN_DESC)
        dyn_tab = None if dyn_tab is None else np.ascontiguousarray(dyn_tab, dtype=np.uint32)
        payload = None if payload is None else np.ascontiguousarray(payload, dtype=np.uint8)
        rc = self.lib.h264gpu_compose_host(_ptr(desc), n, _ptr(hints), len(hints),
                                           _ptr(dyn), 0 if dyn is None else len(dyn),
                                           _ptr(dyn_tab), 0 if dyn_tab is None else len(dyn_tab),
                                           _ptr(payload), 0 if payload is None else payload.size,
                                           _ptr(out), out.size, _ptr(sizes))
        self.check(rc)
        return [out[int(d["out_off"]):int(d["out_off"]) + int(s)].copy() for d, s in zip(desc, sizes)], sizes

    def pinned_copy(self, arr):
        """A copy of `arr` in pinned host memory (h264gpu_host_alloc); returns (pointer, numpy view)."""
        arr = np.ascontiguousarray(arr)
        nb = max(arr.nbytes, 16)
        p = self.lib.h264gpu_host_alloc(nb)
        if not p:
            raise GpuError("pinned allocation of %d bytes failed: %s" % (nb, self.err()))
        v = np.ctypeslib.as_array((C.c_uint8 * nb).from_address(p))
        v[:arr.nbytes] = arr.view(np.uint8).reshape(-1)
        return p, v

    def make_host_batch(self, b, pinned=True):
        """workloads.Batch -> (HostBatch, keep-alive list).  pinned=True copies the arrays into pinned memory once."""
        keep = []
        def ptr(a):
            if a is None or a.size == 0:
                return None
            a = np.ascontiguousarray(a)
            if pinned:
                p, v = self.pinned_copy(a); keep.append((p, v)); return p
            keep.append(a); return a.ctypes.data
        hb = HostBatch()
        hb.desc, hb.n_frames, hb.rdesc = ptr(b.desc), b.n_frames, ptr(b.rdesc)
        hb.regions, hb.n_regions = ptr(b.regions), len(b.regions)
        n_of = lambda a: 0 if a is None else len(a)
        hb.dyn_info, hb.n_dyn_info = ptr(b.dyn_info), n_of(b.dyn_info)
        hb.dyn, hb.n_dyn = ptr(b.dyn), n_of(b.dyn)
        hb.dyn_tab, hb.n_tab = ptr(b.dyn_tab), n_of(b.dyn_tab)
        hb.payload, hb.payload_bytes = ptr(b.payload), n_of(b.payload)
        hb.n_hints = int((b.desc["hint_off"] + b.desc["mb_w"].astype(np.int64) * b.desc["mb_h"]).max())
        hb.slots_bytes = int((b.desc["out_off"] + b.desc["out_cap"]).max())
        hb.max_mbs_per_frame = int((b.desc["mb_w"].astype(np.int64) * b.desc["mb_h"]).max())
        return hb, keep

    def compose_regions_host(self, hb, h_stream, stream_cap, h_offsets, h_sizes=None, n_chunks=0):
        """h264gpu_compose_regions_host: host region lists + payload -> packed Annex-B stream in host memory."""
        self.check(self.lib.h264gpu_compose_regions_host(C.byref(hb), h_stream, stream_cap, h_offsets, h_sizes, n_chunks))

    # ---- device-pointer calls --------------------------------------------------------------------
    def compose_batch(self, d_desc, n, d_hints, d_dyn, d_tab, d_payload, d_out, d_sizes, stream=None, n_hints=None):
        """n_hints=None: the hints the device buffer holds (its size / 8) - the call then also hands the TMA unit a tensor
        map of the dense maps (h264gpu_compose_batch_ex); n_hints=0: plain h264gpu_compose_batch."""
        p = lambda b: b.ptr if b is not None else None
        if n_hints is None:
            n_hints = d_hints.nbytes // 8
        if n_hints:
            self.check(self.lib.h264gpu_compose_batch_ex(p(d_desc), n, p(d_hints), n_hints, p(d_dyn), p(d_tab), p(d_payload),
                                                         p(d_out), p(d_sizes), stream))
        else:
            self.check(self.lib.h264gpu_compose_batch(p(d_desc), n, p(d_hints), p(d_dyn), p(d_tab), p(d_payload),
                                                      p(d_out), p(d_sizes), stream))

    def raster_regions(self, d_rdesc, n, max_mbs, d_regions, d_dyn_info, d_hints, stream=None):
        p = lambda b: b.ptr if b is not None else None
        self.check(self.lib.h264gpu_raster_regions(p(d_rdesc), n, max_mbs, p(d_regions), p(d_dyn_info),
                                                   p(d_hints), stream))

    def canonicalize_hints(self, d_hints, n_hints, stream=None):
        self.check(self.lib.h264gpu_canonicalize_hints(d_hints.ptr, n_hints, stream))

    def rewrite_slices(self, desc, src, out_bytes):
        """Host convenience around h264gpu_rewrite_slices: desc = abi.REWRITE_DESC array, src = uint8 array holding the
        source RBSPs.  Returns (list of NAL byte arrays, sizes)."""
        n = len(desc)
        cb = int(self.lib.h264gpu_rewrite_chunk_bytes())
        total = (40 + desc["hdr_nbits"].astype(np.int64) + desc["n_bits"].astype(np.int64) + 7) // 8
        max_chunks = max(1, int(((total + cb - 1) // cb).max())) if n else 1
        d_desc = self.to_device(desc)
        d_src = self.to_device(np.concatenate([src, np.zeros(16, np.uint8)]))
        d_out, d_sizes, d_cnt = self.alloc(out_bytes), self.alloc(4 * max(n, 1)), self.alloc(4 * (1 + max(n, 1) * max_chunks))
        self.check(self.lib.h264gpu_memset(d_out.ptr, 0xAA, out_bytes, None))
        self.check(self.lib.h264gpu_rewrite_slices(d_desc.ptr, n, max_chunks, d_src.ptr, d_out.ptr, d_sizes.ptr, d_cnt.ptr, None))
        self.sync()
        sizes = d_sizes.download(np.uint32, count=n)
        out = d_out.download(np.uint8)
        res = [out[int(d["out_off"]):int(d["out_off"]) + min(int(s), int(d["out_cap"]))].copy() for d, s in zip(desc, sizes)]
        for d, s in zip(desc, sizes):                    # canaries: nothing behind a slice's last byte inside its slot
            o, c = int(d["out_off"]), int(d["out_cap"])
            if int(s) < c and not (out[o + int(s):o + c] == 0xAA).all():
                raise GpuError("h264gpu_rewrite_slices wrote behind the end of a slice")
        for b_ in (d_desc, d_src, d_out, d_sizes, d_cnt):
            b_.free()
        return res, sizes

    def parse_annexb(self, stream_bytes, max_nals):
        """Host convenience around h264gpu_parse_annexb: uint8 stream -> (abi.NAL_UNIT records, count found)."""
        from . import abi
        n = len(stream_bytes)
        d_data = self.to_device(np.concatenate([stream_bytes, np.full(32, 0xFF, np.uint8)]))
        words = int(self.lib.h264gpu_parse_annexb_scratch_words(n, max_nals))
        d_units, d_scr = self.alloc(12 * max(max_nals, 1)), self.alloc(4 * words)
        self.check(self.lib.h264gpu_parse_annexb(d_data.ptr, n, d_units.ptr, max_nals, d_scr.ptr, None))
        self.sync()
        count = int(d_scr.download(np.uint32, count=max_nals + 1)[max_nals]) if n and max_nals else 0
        units = d_units.download(abi.NAL_UNIT, count=min(count, max_nals))
        for b_ in (d_data, d_units, d_scr):
            b_.free()
        return units, count

    def unescape_nals(self, payloads):
        """Host convenience around h264gpu_unescape_nals: list of uint8 EBSP payloads -> list of RBSP arrays."""
        from . import abi
        n = len(payloads)
        desc = np.zeros(n, abi.UNESCAPE_DESC)
        off, parts = 0, []
        for i, p in enumerate(payloads):
            cap = (len(p) + 15) & ~15
            desc[i]["src_off"], desc[i]["src_bytes"], desc[i]["out_off"], desc[i]["out_cap"] = off, len(p), off, max(cap, 16)
            parts.append(np.concatenate([p, np.zeros(max(cap, 16) - len(p), np.uint8)]))
            off += max(cap, 16)
        cb = int(self.lib.h264gpu_unescape_chunk_bytes())
        max_chunks = max(1, max((len(p) + cb - 1) // cb for p in payloads)) if n else 1
        d_desc = self.to_device(desc)
        d_src = self.to_device(np.concatenate(parts + [np.zeros(32, np.uint8)]))
        d_out, d_sizes, d_st = self.alloc(off + 32), self.alloc(4 * max(n, 1)), self.alloc(4 * (1 + max(n, 1) * max_chunks))
        self.check(self.lib.h264gpu_memset(d_out.ptr, 0xAA, off + 32, None))
        self.check(self.lib.h264gpu_memset(d_sizes.ptr, 0, 4 * max(n, 1), None))
        self.check(self.lib.h264gpu_unescape_nals(d_desc.ptr, n, max_chunks, d_src.ptr, d_out.ptr, d_sizes.ptr, d_st.ptr, None))
        self.sync()
        sizes = d_sizes.download(np.uint32, count=n)
        out = d_out.download(np.uint8)
        res = [out[int(d["out_off"]):int(d["out_off"]) + int(s)].copy() for d, s in zip(desc, sizes)]
        for d, s in zip(desc, sizes):                    # canaries: nothing behind a payload's last byte inside its slot
            o, c = int(d["out_off"]), int(d["out_cap"])
            if int(s) < c and not (out[o + int(s):o + c] == 0xAA).all():
                raise GpuError("h264gpu_unescape_nals wrote behind the end of a payload")
        for b_ in (d_desc, d_src, d_out, d_sizes, d_st):
            b_.free()
        return res

    def pack_frames(self, d_desc, d_sizes, n, d_slots, d_packed, d_offsets, stream=None):
        self.check(self.lib.h264gpu_pack_frames(d_desc.ptr, d_sizes.ptr, n, d_slots.ptr, d_packed.ptr,
                                                d_offsets.ptr, stream))

    def launch_count(self):
        return int(self.lib.h264gpu_launch_count())


# ---- reference-shaped C API (include/h264_writer.h, nal.h) as ctypes structs ----------------------

class WaypointInfo(C.Structure):
    _fields_ = [("offset_px", C.c_int), ("long_term_idx", C.c_int), ("valid", C.c_int)]


class ComposerConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("mb_width", C.c_int), ("mb_height", C.c_int),
                ("log2_max_frame_num", C.c_int), ("pic_order_cnt_type", C.c_int),
                ("log2_max_pic_order_cnt_lsb", C.c_int), ("num_ref_idx_l0_default_minus1", C.c_int),
                ("deblocking_filter_control_present_flag", C.c_int), ("frame_num", C.c_int),
                ("idr_pic_id", C.c_int), ("waypoints", WaypointInfo * 8), ("num_waypoints", C.c_int)]


class NALWriter(C.Structure):
    _fields_ = [("output", C.c_void_p), ("output_capacity", C.c_size_t), ("output_pos", C.c_size_t),
                ("rbsp", C.c_void_p), ("rbsp_capacity", C.c_size_t)]


class Composer(C.Structure):
    _fields_ = [("cfg", ComposerConfig), ("parse_cfg", ComposerConfig),
                ("ref_a_rbsp", C.c_void_p), ("ref_a_size", C.c_size_t),
                ("ref_b_rbsp", C.c_void_p), ("ref_b_size", C.c_size_t),
                ("orig_sps", C.c_void_p), ("orig_sps_size", C.c_size_t),
                ("orig_pps", C.c_void_p), ("orig_pps_size", C.c_size_t),
                ("nw", NALWriter), ("output_buffer", C.c_void_p), ("output_capacity", C.c_size_t),
                ("rbsp_temp", C.c_void_p), ("rbsp_capacity", C.c_size_t), ("frames_written", C.c_int)]


class ComposerMotionRegion(C.Structure):
    _fields_ = [("mb_x0", C.c_int16), ("mb_y0", C.c_int16), ("mb_x1", C.c_int16), ("mb_y1", C.c_int16),
                ("mv_x_px", C.c_int16), ("mv_y_px", C.c_int16), ("ref_idx", C.c_uint8)]


class ComposerMbInfo(C.Structure):
    _fields_ = [("mv_x_qpel", C.c_int16), ("mv_y_qpel", C.c_int16), ("ref_idx", C.c_uint8), ("intra", C.c_uint8)]


class ComposerDynRegion(C.Structure):
    _fields_ = [("mb_x0", C.c_int16), ("mb_y0", C.c_int16), ("mb_x1", C.c_int16), ("mb_y1", C.c_int16),
                ("payload", C.c_void_p), ("payload_bytes", C.c_size_t), ("mb_bit_off", C.c_void_p), ("mb_info", C.c_void_p)]


class ComposerFrameHints(C.Structure):
    _fields_ = [("offset_px", C.c_int), ("regions", C.c_void_p), ("n_regions", C.c_int), ("dyn", C.c_void_p)]
