"""
TEST INFRASTRUCTURE: a minimal CAVLC P-slice motion decoder, written from the H.264 specification text and from nothing in
this repository's kernels or oracle - it is the independent party that certifies the SPEC modes (H264GPU_SKIP_SPEC |
H264GPU_PRED_SPEC), for which the reference has no counterpart (no ffmpeg in the image).

What it understands is exactly what the composer writes: a slice of P_L0_16x16 macroblocks without residual
(coded_block_pattern 0), P_Skip macroblocks, and opaque pre-encoded macroblocks whose bit lengths and neighbour-visible
motion the caller supplies (the splice contract: they are not re-parsed, docs/MASTER_DESIGN.md:142-171).

  7.3.4    slice_data():  mb_skip_run ue(v) | macroblock_layer()            (CAVLC: entropy_coding_mode_flag = 0)
  7.3.5    macroblock_layer(): mb_type ue(v) = 0 (P_L0_16x16) | ref_idx_l0 te(v) | mvd_l0 se(v) x 2 | coded_block_pattern me(v)
  9.1      ue(v) / se(v) / te(v):  te(v) with range 1 is one inverted bit, else ue(v)
  8.4.1.1  P_Skip motion:  mv = 0 when mbAddrA or mbAddrB is not available, or A / B has refIdx 0 and mv 0; else the predictor
  8.4.1.3  median luma motion vector prediction, 8.4.1.3.2 neighbouring partitions (C falls back to D when not available)
  7.4.1.1  emulation prevention: 00 00 03 -> 00 00 inside the NAL unit
"""
import numpy as np


class Bits:
    def __init__(self, rbsp):
        self.bits = np.unpackbits(np.frombuffer(bytes(rbsp), np.uint8))
        self.pos = 0
        ones = np.nonzero(self.bits)[0]
        self.last_one = int(ones[-1]) if len(ones) else -1      # the rbsp_stop_one_bit

    def u(self, n):
        v = 0
        for b in self.bits[self.pos:self.pos + n]:
            v = (v << 1) | int(b)
        if self.pos + n > len(self.bits):
            raise ValueError("read past the end of the slice")
        self.pos += n
        return v

    def ue(self):
        z = 0
        while self.u(1) == 0:
            z += 1
            if z > 32:
                raise ValueError("not an Exp-Golomb code")
        return (1 << z) - 1 + (self.u(z) if z else 0)

    def se(self):
        k = self.ue()
        return (k + 1) // 2 if k & 1 else -(k // 2)

    def te(self, rng):
        if rng <= 0:
            return 0
        return 1 - self.u(1) if rng == 1 else self.ue()

    def more_rbsp_data(self):
        """7.2: there is more data when something other than the rbsp_trailing_bits (a 1 and zeros) follows."""
        return self.pos < self.last_one


def unescape(nal_payload):
    """7.4.1.1: drop emulation_prevention_three_byte (a 03 behind two zero bytes)."""
    out, zeros = bytearray(), 0
    for b in bytes(nal_payload):
        if zeros >= 2 and b == 3:
            zeros = 0
            continue
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


def _median(a, b, c):
    return a + b + c - min(a, b, c) - max(a, b, c)


class _Nbr:
    """What 8.4.1.3.2 yields for a neighbouring 16x16 partition: availability, refIdx (-1: not available or intra), mv."""
    __slots__ = ("avail", "ref", "mv")

    def __init__(self, avail=False, ref=-1, mv=(0, 0)):
        self.avail, self.ref, self.mv = avail, ref, mv


def _predict(A, B, C, ref):
    """8.4.1.3 for a 16x16 partition (the directional rules of 16x8 / 8x16 do not apply)."""
    if not B.avail and not C.avail and A.avail:
        B, C = A, A
    m = [n for n in (A, B, C) if n.ref == ref]
    if len(m) == 1:
        return m[0].mv
    return (_median(A.mv[0], B.mv[0], C.mv[0]), _median(A.mv[1], B.mv[1], C.mv[1]))


def decode_p_slice(nal, mb_w, mb_h, num_refs, hdr_nbits, opaque=None):
    """nal = Annex-B NAL unit (start code, header byte, EBSP).  opaque[mb] = (n_bits, ref, (mvx, mvy)) for pre-encoded
    macroblocks (ref -1: intra).  Returns per macroblock (skipped, ref, (mvx, mvy)) in raster order; pre-encoded ones
    come back as (False, ref, mv) from `opaque`."""
    nal = bytes(nal)
    assert nal[:4] == b"\x00\x00\x00\x01" and (nal[4] & 0x1F) == 1, "not a non-IDR slice NAL unit"
    br = Bits(unescape(nal[5:]))
    br.u(hdr_nbits)                                   # slice_header(): written (and checked) on the host side
    n = mb_w * mb_h
    out = [None] * n
    opaque = opaque or {}

    def nbr(x, y):
        if x < 0 or y < 0 or x >= mb_w:
            return _Nbr()
        r = out[y * mb_w + x]
        return _Nbr(True, r[1], r[2]) if r is not None else _Nbr()

    mb = 0
    more = True
    while more:
        run = br.ue()                                 # mb_skip_run
        for _ in range(run):
            if mb >= n:
                raise ValueError("mb_skip_run runs past the picture")
            x, y = mb % mb_w, mb // mb_w
            A, B = nbr(x - 1, y), nbr(x, y - 1)
            C = nbr(x + 1, y - 1)
            if not C.avail:
                C = nbr(x - 1, y - 1)
            if (not A.avail or not B.avail or (A.ref == 0 and A.mv == (0, 0)) or (B.ref == 0 and B.mv == (0, 0))):
                mv = (0, 0)
            else:
                mv = _predict(A, B, C, 0)
            out[mb] = (True, 0, mv)
            mb += 1
        more = br.more_rbsp_data()
        if mb >= n:
            break
        if not more:
            raise ValueError("slice data ends at macroblock %d of %d" % (mb, n))
        x, y = mb % mb_w, mb // mb_w
        if mb in opaque:                              # pre-encoded macroblock: its bits are the dynamic encoder's business
            nb, ref, mv = opaque[mb]
            br.u(0)
            br.pos += nb
            out[mb] = (False, ref, mv)
        else:
            assert br.ue() == 0, "mb_type must be P_L0_16x16"
            ref = br.te(num_refs - 1)
            dx, dy = br.se(), br.se()
            assert br.ue() == 0, "coded_block_pattern must be 0"
            A, B = nbr(x - 1, y), nbr(x, y - 1)
            C = nbr(x + 1, y - 1)
            if not C.avail:
                C = nbr(x - 1, y - 1)
            p = _predict(A, B, C, ref)
            out[mb] = (False, ref, (p[0] + dx, p[1] + dy))
        mb += 1
        more = br.more_rbsp_data()
    assert mb == n, "decoded %d of %d macroblocks" % (mb, n)
    # rbsp_slice_trailing_bits: a one, then zeros to the byte boundary
    rest = br.bits[br.pos:]
    assert len(rest) >= 1 and rest[0] == 1 and not rest[1:].any(), "bad rbsp_trailing_bits"
    return out
