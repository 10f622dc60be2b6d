"""numpy mirrors of the PODs in include/composer_gpu.h (little-endian, packed as the C compiler lays them out)."""
import numpy as np

MB_INTER, MB_DYN_INTER, MB_DYN_INTRA, REGION_DYN = 0, 1, 2, 3
SKIP_OFF, SKIP_SPEC, PRED_REF, PRED_SPEC = 0, 1, 0, 2
NO_DYN = 0xFFFFFFFF

HINT = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("ref", "u1"), ("kind", "u1"), ("dyn", "<u2")])
FRAME_DESC = np.dtype([
    ("hdr_bits", "<u8", (4,)), ("hint_off", "<u8"), ("out_off", "<u8"),
    ("out_cap", "<u4"), ("dyn_id", "<u4"), ("mb_w", "<u2"), ("mb_h", "<u2"),
    ("hdr_nbits", "u1"), ("num_refs", "u1"), ("flags", "u1"), ("nal_header", "u1")])
DYN_DESC = np.dtype([("payload_off", "<u8"), ("table_off", "<u8"), ("n_mbs", "<u4"),
                     ("payload_bytes", "<u4"), ("reserved", "<u8")])
REGION = np.dtype([("x0", "<i2"), ("y0", "<i2"), ("x1", "<i2"), ("y1", "<i2"),
                   ("mvx", "<i2"), ("mvy", "<i2"), ("ref", "u1"), ("kind", "u1"), ("dyn_base", "<u2")])
RASTER_DESC = np.dtype([("region_first", "<u4"), ("region_count", "<u4"), ("hint_off", "<u8"),
                        ("mb_w", "<u2"), ("mb_h", "<u2"), ("dyn_info_off", "<u4")])

REWRITE_DESC = np.dtype([("hdr_bits", "<u8", (4,)), ("src_off", "<u8"), ("src_bytes", "<u8"), ("src_start_bit", "<u8"),
                         ("n_bits", "<u8"), ("out_off", "<u8"), ("out_cap", "<u4"), ("hdr_nbits", "<u2"),
                         ("nal_header", "u1"), ("reserved", "u1")])

UNESCAPE_DESC = np.dtype([("src_off", "<u8"), ("src_bytes", "<u8"), ("out_off", "<u8"), ("out_cap", "<u4"), ("reserved", "<u4")])

NAL_UNIT = np.dtype([("offset", "<u4"), ("size", "<u4"), ("nal_ref_idc", "u1"), ("nal_unit_type", "u1"), ("reserved", "u1", (2,))])

assert REWRITE_DESC.itemsize == 80 and UNESCAPE_DESC.itemsize == 32 and NAL_UNIT.itemsize == 12
assert HINT.itemsize == 8 and FRAME_DESC.itemsize == 64 and DYN_DESC.itemsize == 32
assert REGION.itemsize == 16 and RASTER_DESC.itemsize == 24


def hdr_words(bits: str) -> np.ndarray:
    """'0101..' -> 4 x u64, MSB first, zero padded (layout of H264GpuFrameDesc.hdr_bits)."""
    assert len(bits) <= 255
    padded = bits.ljust(256, "0")
    return np.array([int(padded[64 * k:64 * k + 64], 2) for k in range(4)], dtype=np.uint64)
