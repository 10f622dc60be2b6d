/*
 * mb_codec.cuh - per-macroblock arithmetic of the P-slice composer, shared by
 * the sm_100a kernels (compose_kernels.cu) and by the CPU emulation harness in
 * tests/emu (which compiles this header with g++ to check the bit twiddling
 * against the oracle without a GPU).  Everything here is a pure function.
 *
 * Reference behaviour being reproduced (paths under /root/reference):
 *   median3 / get_mv_prediction   src/h264_writer.c:362-432
 *   write_p16x16_mb               src/h264_writer.c:434-453
 *   bitwriter_write_ue / _se      src/bitwriter.c:50-101
 *   rbsp_to_ebsp                  src/nal.c:24-50
 */
#ifndef H264SCROLL_MB_CODEC_CUH
#define H264SCROLL_MB_CODEC_CUH

#include <stdint.h>
#include "composer_gpu.h"

#if defined(__CUDACC__)
#define HSD __host__ __device__ __forceinline__
#else
#define HSD static inline
#endif

namespace h264scroll {

typedef uint32_t u32;
typedef uint64_t u64;

/* ---------------------------------------------------------------- bit ops */

HSD int clz32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

/* high 32 bits of (hi:lo) << s, 0 <= s <= 31 */
HSD u32 fshl(u32 hi, u32 lo, int s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

HSD u32 bswap32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}

/* 1 where a byte of v is zero (exact) */
HSD u32 zero_byte_mask(u32 v) {
    return ~((((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) | 0x7F7F7F7Fu);
}

/* ------------------------------------------------------------ Exp-Golomb */

/* ue(v) = (v+1) written in 2*floor(log2(v+1))+1 bits (src/bitwriter.c:50-74) */
HSD int ue_nbits(u32 v) { return 2 * (31 - clz32(v + 1)) + 1; }

/* se(v) code number: v>0 -> 2v-1, else -2v (src/bitwriter.c:91-101) */
HSD u32 se_codenum(int v) { return v > 0 ? (u32)(2 * v - 1) : (u32)(-2 * v); }

/* -------------------------------------------------------- MB table entry
 *
 * Every macroblock becomes one 16-byte entry plus a bit offset:
 *   w0            mb_skip_run prefix, left aligned (ue(run), <= 29 bits)
 *   inline MB :   w1,w2,w3   body bits left aligned (mb_type|ref|mvd|cbp, <= 85 bits);
 *                 the low 11 bits of w3 are never body bits
 *   spliced MB:   w1 = payload bit offset, w2 = payload bit length
 *   w3 low bits   [5:0] prefix length, [6] spliced flag
 */
struct alignas(16) MbEntry { u32 w0, w1, w2, w3; };

#define ENT_ULEN_MASK 0x3Fu
#define ENT_DYN       0x40u
#define ENT_META_MASK 0x7FFu

/*
 * The bit-offset word of an entry carries its kind in the two top bits
 * (slice bit offsets stay below 2^30):
 *   KIND_SHORT  prefix+body fit one word: w0 holds the whole code, left aligned
 *   KIND_LONG   inline MB longer than 32 bits: w0 prefix, w1..w3 body
 *   KIND_DYN    spliced MB: w0 prefix, w1 payload bit offset, w2 payload bit length
 */
#define KIND_SHORT 0u
#define KIND_LONG  1u
#define KIND_DYN   2u
#define OFF_KIND_SHIFT 30
#define OFF_MASK 0x3FFFFFFFu

/* append the low n (1..33) bits of v to the right-aligned accumulator hi:lo */
HSD void acc_push(u64 &hi, u64 &lo, u64 v, int n) {
    hi = (hi << n) | (lo >> (64 - n));
    lo = (lo << n) | v;
}

/* neighbour-visible motion of a hint (intra: ref 255, mv 0; src/h264_writer.c:636-640) */
struct Nbr { int mvx, mvy, ref; };

HSD Nbr nbr_from_hint(u32 lo, u32 hi) {
    Nbr n;
    u32 kind = (hi >> 8) & 0xFF;
    if (kind == H264GPU_MB_DYN_INTRA) { n.mvx = 0; n.mvy = 0; n.ref = 255; }
    else {
        n.mvx = (int)(int16_t)(lo & 0xFFFF);
        n.mvy = (int)(int16_t)(lo >> 16);
        n.ref = (int)(hi & 0xFF);
    }
    return n;
}

HSD int imin_(int a, int b) { return a < b ? a : b; }
HSD int imax_(int a, int b) { return a > b ? a : b; }

/*
 * MV predictor.  avA/avB/avC are availability flags; unavailable neighbours
 * must be passed with mv 0.  spec = 0 reproduces the reference including
 * median3 == min(max(a,b),c) and the single-neighbour rule; spec = 1 is
 * H.264 8.4.1.3.
 */
HSD void predict_mv(int spec, Nbr a, int avA, Nbr b, int avB, Nbr c, int avC, int ref,
                    int &px, int &py) {
    if (spec && !avB && !avC && avA) { b = a; c = a; avB = 1; avC = 1; }
    int ma = avA && a.ref == ref, mb = avB && b.ref == ref, mc = avC && c.ref == ref;
    int nav = avA + avB + avC, nm = ma + mb + mc;
    if (!spec && nav <= 1) {
        /* zero or one neighbour: its MV when the ref matches, else 0 (:404-417) */
        px = ma ? a.mvx : (mb ? b.mvx : (mc ? c.mvx : 0));
        py = ma ? a.mvy : (mb ? b.mvy : (mc ? c.mvy : 0));
        return;
    }
    if (nm == 1) {
        px = ma ? a.mvx : (mb ? b.mvx : c.mvx);
        py = ma ? a.mvy : (mb ? b.mvy : c.mvy);
        return;
    }
    int hx = imin_(imax_(a.mvx, b.mvx), c.mvx), hy = imin_(imax_(a.mvy, b.mvy), c.mvy);
    if (spec) {
        hx = imax_(imin_(a.mvx, b.mvx), hx);
        hy = imax_(imin_(a.mvy, b.mvy), hy);
    }
    px = hx; py = hy;
}

/* H.264 8.4.1.1: inferred P_Skip MV */
HSD void skip_infer_mv(int spec, Nbr a, int avA, Nbr b, int avB, Nbr c, int avC, int &sx, int &sy) {
    if (!avA || !avB || (a.ref == 0 && a.mvx == 0 && a.mvy == 0) ||
        (b.ref == 0 && b.mvx == 0 && b.mvy == 0)) { sx = 0; sy = 0; return; }
    predict_mv(spec, a, avA, b, avB, c, avC, 0, sx, sy);
}

/*
 * Body of a P_L0_16x16 macroblock: '1' | te(ref) | se(mvdx) | se(mvdy) | '1'
 * (src/h264_writer.c:434-453).  Returns the bit count; body bits left aligned in
 * e.w1..w3 (low 11 bits of w3 cleared).
 */
HSD int encode_p16x16_body(int ref, int mvdx, int mvdy, int num_refs, MbEntry &e) {
    {   /* common case: the whole body fits one word */
        const u32 kx = se_codenum(mvdx), ky = se_codenum(mvdy);
        const int nx = ue_nbits(kx), ny = ue_nbits(ky);
        const int rl = num_refs == 2 ? 1 : (num_refs > 2 ? ue_nbits((u32)ref) : 0);
        const int total = 2 + rl + nx + ny;
        if (total <= 32) {
            const u32 rc = num_refs == 2 ? (u32)(1 - (ref & 1)) : (u32)ref + 1u;
            u32 v = 1u;
            v = (v << rl) | (rl ? rc : 0u);
            v = (v << nx) | (kx + 1u);
            v = (v << ny) | (ky + 1u);
            v = (v << 1) | 1u;
            e.w1 = v << (32 - total);
            e.w2 = 0; e.w3 = 0;
            return total;
        }
    }
    u64 hi = 0, lo = 1;          /* mb_type ue(0) */
    int len = 1;
    if (num_refs == 2) { acc_push(hi, lo, (u64)(1 - (ref & 1)), 1); len += 1; }
    else if (num_refs > 2) { int n = ue_nbits((u32)ref); acc_push(hi, lo, (u64)ref + 1, n); len += n; }
    u32 kx = se_codenum(mvdx), ky = se_codenum(mvdy);
    int nx = ue_nbits(kx), ny = ue_nbits(ky);
    acc_push(hi, lo, (u64)kx + 1, nx);
    acc_push(hi, lo, (u64)ky + 1, ny);
    acc_push(hi, lo, 1, 1);      /* coded_block_pattern ue(0) */
    len += nx + ny + 1;
    /* left-align the len-bit value in 96 bits */
    int sh = 96 - len;           /* 11 .. 93 */
    u64 h2, l2;
    if (sh >= 64) { h2 = lo << (sh - 64); l2 = 0; }
    else { h2 = (hi << sh) | (lo >> (64 - sh)); l2 = lo << sh; }
    /* value now occupies bits [95:0] of h2:l2 with h2 holding bits [95:64] in its low half */
    e.w1 = (u32)h2;
    e.w2 = (u32)(l2 >> 32);
    e.w3 = (u32)l2 & ~ENT_META_MASK;
    return len;
}

/*
 * Body of an inline MB whose mvd is (0,0): '1' te(ref) '1' '1' '1', left aligned.
 * This is what every MB in the interior of a uniformly moving region codes.
 */
HSD u32 zero_mvd_body(int ref, int num_refs, int &blen) {
    if (num_refs == 2) { blen = 5; return (0x17u | ((u32)(1 - (ref & 1)) << 3)) << 27; }
    if (num_refs < 2) { blen = 4; return 0xFu << 28; }
    const int n = ue_nbits((u32)ref);
    blen = n + 4;
    u32 v = (1u << n) | ((u32)ref + 1u);
    v = (v << 3) | 7u;
    return v << (32 - blen);
}

/* 4-bit mask of the zero bytes of a word: bit k = byte k, byte 0 being the first (most significant) */
HSD u32 zero_nibble(u32 w) {
    const u32 z = zero_byte_mask(w);
    return ((z >> 31) & 1u) | ((z >> 22) & 2u) | ((z >> 13) & 4u) | ((z >> 4) & 8u);
}

/* mb_skip_run prefix: ue(run) left aligned in a word + its length */
HSD void encode_skip_prefix(u32 run, u32 &bits, int &n) {
    n = ue_nbits(run);
    bits = (run + 1) << (32 - n);
}

/* ------------------------------------------------------- payload access */

/* 32 payload bits starting at bit position p (MSB first), words are read big-endian */
HSD u32 payload_bits32(const u32 *payload_words, u32 n_words, u32 p) {
    u32 wi = p >> 5;
    u32 a = wi < n_words ? bswap32(payload_words[wi]) : 0u;
    u32 b = wi + 1 < n_words ? bswap32(payload_words[wi + 1]) : 0u;
    return fshl(a, b, (int)(p & 31));
}

/*
 * Bits that MB entry `e`, whose first bit sits at frame bit `o`, contributes to
 * the 32-bit output word starting at frame bit `base`.
 */
HSD u32 entry_contrib(const MbEntry &e, u32 o, u32 base, const u32 *payload_words, u32 n_payload_words) {
    int q = (int)(base - o);             /* position of the word inside the MB string */
    int ulen = (int)(e.w3 & ENT_ULEN_MASK);
    u32 acc = 0;
    if (q < 0) { if (q > -32) acc = e.w0 >> (-q); }
    else if (q < 32) acc = e.w0 << q;
    int q2 = q - ulen;                   /* position inside the body */
    if (!(e.w3 & ENT_DYN)) {
        u32 b0 = e.w1, b1 = e.w2, b2 = e.w3 & ~ENT_META_MASK;
        if (q2 < 0) { if (q2 > -32) acc |= b0 >> (-q2); }
        else if (q2 < 96) {
            int ws = q2 >> 5, sh = q2 & 31;
            u32 x = ws == 0 ? b0 : (ws == 1 ? b1 : b2);
            u32 y = ws == 0 ? b1 : (ws == 1 ? b2 : 0u);
            acc |= fshl(x, y, sh);
        }
    } else {
        u32 pb = e.w1, L = e.w2;
        if (q2 < 0) {
            if (q2 > -32 && L) {
                u32 v = payload_bits32(payload_words, n_payload_words, pb);
                if (L < 32) v &= ~(0xFFFFFFFFu >> L);
                acc |= v >> (-q2);
            }
        } else if ((u32)q2 < L) {
            u32 v = payload_bits32(payload_words, n_payload_words, pb + (u32)q2);
            u32 rem = L - (u32)q2;
            if (rem < 32) v &= ~(0xFFFFFFFFu >> rem);
            acc |= v;
        }
    }
    return acc;
}

/* contribution of a KIND_SHORT entry (whole code in one left-aligned word) */
HSD u32 short_contrib(u32 code, u32 o, u32 base) {
    int q = (int)(base - o);
    if (q < 0) return q > -32 ? code >> (-q) : 0u;
    return q < 32 ? code << q : 0u;
}

/*
 * Attach the mb_skip_run prefix (pbits left aligned, pn bits) to an entry whose
 * body (blen bits) sits left aligned in w1..w3.  Returns the kind.
 */
HSD u32 finish_inline_entry(MbEntry &e, u32 pbits, int pn, int blen) {
    if (pn + blen <= 32) {
        e.w0 = pbits | (pn < 32 ? e.w1 >> pn : 0u);
        return KIND_SHORT;
    }
    e.w0 = pbits;
    e.w3 = (e.w3 & ~ENT_META_MASK) | (u32)pn;
    return KIND_LONG;
}

/* --------------------------------------------------------------- header */

/* 32 bits at stream bit position p of: 00 00 00 01 | nal_header | slice header bits */
HSD u32 header_bits32(const u64 hdr[4], u32 nal_header, u32 p) {
    /* string = 40-bit prefix followed by 256 header bits */
    u32 acc = 0;
    /* prefix words: bits 0..31 = 0x00000001, bits 32..39 = nal_header */
    if (p < 32) acc |= 0x00000001u << p;                /* p == 0 keeps the word, else shifts the 1 out */
    if (p < 40) {
        int d = (int)p - 32;                            /* where the header byte starts relative to p */
        u32 nh = nal_header << 24;
        if (d >= 0) acc |= nh << d;
        else if (d > -32) acc |= nh >> (-d);
    }
    /* slice header bits start at stream bit 40 */
    int q = (int)p - 40;
    for (int k = 0; k < 4; k++) {
        /* hdr[k] covers header bits [64k, 64k+64) */
        int rel = q - 64 * k;                           /* offset of p inside hdr[k] */
        if (rel <= -32 || rel >= 64) continue;
        u64 w = hdr[k];
        u32 piece;
        if (rel < 0) piece = (u32)(w >> 32) >> (-rel);
        else if (rel <= 32) piece = (u32)(w >> (32 - rel));
        else piece = (u32)(w << (rel - 32));
        acc |= piece;
    }
    return acc;
}

/* --------------------------------------------------- emulation prevention
 *
 * rbsp_to_ebsp (src/nal.c:24-50) in closed form: with r = number of zero bytes
 * immediately before RBSP byte p (continuing the automaton state zc_in carried
 * into this chunk), 0x03 is inserted before p iff byte(p) <= 3, r >= 2, r even.
 *
 * words[] holds the chunk MSB-first (stream byte 4*j+k is bits [31-8k:24-8k] of
 * words[j]); chunk_byte0 is the stream byte index of words[0]; stream bytes
 * below 5 are the start code and NAL header and never take part.
 */
HSD u32 chunk_byte(const u32 *words, int b) { return (words[b >> 2] >> (24 - 8 * (b & 3))) & 0xFFu; }

HSD u32 epb_insert_mask(const u32 *words, int j, u32 chunk_byte0, int zc_in, int valid_bytes) {
    u32 mask = 0;
    for (int k = 0; k < valid_bytes; k++) {
        int b = 4 * j + k;
        if (chunk_byte0 + (u32)b < 5) continue;
        if (chunk_byte(words, b) > 3) continue;
        int r = 0, c = b - 1;
        while (c >= 0 && chunk_byte0 + (u32)c >= 5 && chunk_byte(words, c) == 0) { r++; c--; }
        if (c < 0) r += zc_in;
        if (r >= 2 && !(r & 1)) mask |= 1u << k;
    }
    return mask;
}

/*
 * Scan form of the same rule, one output word at a time.  b0 = chunk-relative
 * index of the word's first byte, prev_nz = chunk-relative index of the last
 * non-zero byte before it (may be negative: -1 - zc_in continues the carried
 * automaton state).  Returns the 4-bit insert mask and updates prev_nz.
 */
HSD u32 epb_word_mask(u32 w, int vb, int b0, u32 chunk_byte0, int &prev_nz) {
    u32 mask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        if (k < vb) {
            const u32 v = (w >> (24 - 8 * k)) & 0xFFu;
            const int b = b0 + k;
            const int r = b - 1 - prev_nz;
            if (v <= 3u && r >= 2 && !(r & 1) && chunk_byte0 + (u32)b >= 5u) mask |= 1u << k;
            if (v != 0u) prev_nz = b;
        }
    }
    return mask;
}

/* chunk-relative index of the last non-zero byte of a word (first vb bytes), or EPB_NONE */
#define EPB_NONE (-0x40000000)
HSD int word_last_nonzero(u32 w, int vb, int b0) {
    int r = EPB_NONE;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++)
        if (k < vb && ((w >> (24 - 8 * k)) & 0xFFu) != 0u) r = b0 + k;
    return r;
}

/* automaton state after a chunk whose emitted bytes are words[0 .. nbytes) */
HSD int epb_state_after(const u32 *words, int nbytes, u32 chunk_byte0, int zc_in) {
    int r = 0, c = nbytes - 1;
    while (c >= 0 && chunk_byte0 + (u32)c >= 5 && chunk_byte(words, c) == 0) { r++; c--; }
    if (c < 0) r += zc_in;
    return r == 0 ? 0 : ((r & 1) ? 1 : 2);
}


/* ======================================================================
 * Register pipeline (compose kernel v3): every thread owns MPT consecutive
 * macroblocks, keeps their codes in registers and appends them to the slice
 * with a sequential bit writer.  Pure functions, shared with tests/emu.
 * ==================================================================== */

/* packed int16x2 max / min (mv_x in the low half, mv_y in the high half) */
HSD u32 vmax2(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    int16_t al = (int16_t)(a & 0xFFFF), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xFFFF), bh = (int16_t)(b >> 16);
    return (u32)(uint16_t)(al > bl ? al : bl) | ((u32)(uint16_t)(ah > bh ? ah : bh) << 16);
#endif
}
HSD u32 vmin2(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __vmins2(a, b);
#else
    int16_t al = (int16_t)(a & 0xFFFF), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xFFFF), bh = (int16_t)(b >> 16);
    return (u32)(uint16_t)(al < bl ? al : bl) | ((u32)(uint16_t)(ah < bh ? ah : bh) << 16);
#endif
}

/* per-MB register record: `meta` beside a 64-bit `code` */
#define META_LEN_MASK  0x0000FFFFu   /* bits this item appends through the bit writer (prefix + inline body) */
#define META_DYN       0x00010000u   /* spliced MB: code = (dyn index << 48) | prefix value; the payload follows */
#define META_LONG      0x00020000u   /* inline MB that may not fit 64 bits with its prefix:
                                        code = mvd_x (18 bits) | mvd_y << 18 | prefix value << 36, rebuilt when written */
#define META_CODED     0x00080000u   /* not P_Skip */
#define META_REF_SHIFT 24
#define PREFIX_MAX_OFF   1           /* mb_skip_run is always ue(0) with skip OFF (src/h264_writer.c:629-631) */
#define PREFIX_MAX_SPEC 41           /* ue(run), run < 2^20 */

/* neighbour-visible motion of a hint: packed mv + ref key (0x100 = unavailable, never matches) */
HSD void nbr_view(u32 lo, u32 hi, int av, u32 &mv, u32 &ref) {
    const bool intra = ((hi >> 8) & 0xFFu) == H264GPU_MB_DYN_INTRA;
    mv = (av && !intra) ? lo : 0u;
    ref = av ? (intra ? 255u : (hi & 0xFFu)) : 0x100u;
}

/*
 * MV predictor on packed int16x2 values (src/h264_writer.c:369-432; spec = H.264 8.4.1.3).
 * Unavailable neighbours carry mv 0 / ref 0x100.  With zero or one neighbour available and no
 * ref match the reference returns 0, which is what min(max(a,b),c) gives with the zeros in place.
 */
HSD u32 predict_packed(int spec, u32 mva, u32 ra, u32 mvb, u32 rb, u32 mvc, u32 rc, int only_a, u32 ref) {
    if (spec && only_a) { mvb = mva; mvc = mva; rb = ra; rc = ra; }
    const int ma = ra == ref, mb = rb == ref, mc = rc == ref;
    u32 med = vmin2(vmax2(mva, mvb), mvc);                 /* the reference's median3 (:362-367) */
    if (spec) med = vmax2(vmin2(mva, mvb), med);           /* the true median */
    const u32 single = ma ? mva : (mb ? mvb : mvc);
    return (ma + mb + mc == 1) ? single : med;
}

/* se(v): value written (code number + 1) and its bit count */
HSD u32 se_codeval(int v, int &n) {
    const u32 t = (u32)v << 1;
    const u32 cv = v > 0 ? t : 1u - t;
    n = 63 - 2 * clz32(cv);
    return cv;
}

/* body '1' te(ref) se(dx) se(dy) '1' right aligned in 64 bits; returns its length (body valid when <= 64) */
HSD int inline_body64(int ref, int dx, int dy, int num_refs, u64 &body) {
    int nx, ny, rl = 0;
    u32 rc = 0;
    const u32 cx = se_codeval(dx, nx), cy = se_codeval(dy, ny);
    if (num_refs == 2) { rc = (u32)(~ref) & 1u; rl = 1; }
    else if (num_refs > 2) { rc = (u32)ref + 1u; rl = 63 - 2 * clz32(rc); }
    const int blen = 2 + rl + nx + ny;
    if (blen <= 32) {
        u32 v = (1u << rl) | rc;
        v = (v << nx) | cx;
        v = (v << ny) | cy;
        body = (u64)((v << 1) | 1u);
    } else if (blen <= 64) {
        u64 v = ((u64)1 << rl) | rc;
        v = (v << nx) | cx;
        v = (v << ny) | cy;
        body = (v << 1) | 1u;
    } else body = 0;
    return blen;
}

/*
 * One inter macroblock: neighbours -> predictor -> P_Skip test -> mvd -> body.
 * lo/hi words of the own hint and of A, B, C(D); av* availability.
 * Returns meta (length, flags, ref) and the 64-bit code (no mb_skip_run prefix yet).
 * Bodies longer than 64 - prefix_max bits become META_LONG records.
 */
HSD u32 decide_inter(u32 hlo, u32 hhi, u32 alo, u32 ahi, int avA, u32 blo, u32 bhi, int avB,
                     u32 clo, u32 chi, int avC, int num_refs, int pred_spec, int skip_spec, u64 &code) {
    const u32 ref = hhi & 0xFFu;
    u32 mva, ra, mvb, rb, mvc, rc;
    nbr_view(alo, ahi, avA, mva, ra);
    nbr_view(blo, bhi, avB, mvb, rb);
    nbr_view(clo, chi, avC, mvc, rc);
    const u32 pred = predict_packed(pred_spec, mva, ra, mvb, rb, mvc, rc, avA && !avB && !avC, ref);
    if (skip_spec && ref == 0u) {
        /* H.264 8.4.1.1: the inferred motion of P_Skip */
        const bool zero = !avA || !avB || (ra == 0u && mva == 0u) || (rb == 0u && mvb == 0u);
        if (hlo == (zero ? 0u : pred)) { code = 0; return 0u; }
    }
    const int dx = (int)(int16_t)(hlo & 0xFFFFu) - (int)(int16_t)(pred & 0xFFFFu);
    const int dy = (int)(int16_t)(hlo >> 16) - (int)(int16_t)(pred >> 16);
    const int blen = inline_body64((int)ref, dx, dy, num_refs, code);
    u32 meta = (u32)blen | META_CODED | (ref << META_REF_SHIFT);
    if (blen > 64 - (skip_spec ? PREFIX_MAX_SPEC : PREFIX_MAX_OFF)) {
        code = (u64)((u32)dx & 0x3FFFFu) | ((u64)((u32)dy & 0x3FFFFu) << 18);
        meta |= META_LONG;
    }
    return meta;
}

/*
 * The same for a macroblock whose neighbourhood is uniform (A, B and C available and equal to
 * the MB's own hint): every predictor rule yields the MB's own MV, so mvd = 0; with ref 0 that
 * is also exactly the inferred P_Skip motion.
 */
HSD u32 decide_uniform(u32 hhi, int num_refs, int skip_spec, u64 &code) {
    const u32 ref = hhi & 0xFFu;
    if (skip_spec && ref == 0u) { code = 0; return 0u; }
    int blen;
    code = (u64)(zero_mvd_body((int)ref, num_refs, blen) >> (32 - blen));
    return (u32)blen | META_CODED | (ref << META_REF_SHIFT);
}

HSD int long_dx(u64 code) { return ((int)((u32)code << 14)) >> 14; }
HSD int long_dy(u64 code) { return ((int)((u32)(code >> 18) << 14)) >> 14; }
HSD u32 long_prefix(u64 code) { return (u32)(code >> 36); }
HSD u32 dyn_prefix(u64 code) { return (u32)code; }
HSD u32 dyn_index(u64 code) { return (u32)(code >> 48); }

/* attach the mb_skip_run prefix ue(run) (value pre = run + 1 < 2^21, in pn bits) to a coded record */
HSD void attach_prefix(u64 &code, u32 &meta, u32 pre, int pn) {
    if (meta & META_LONG) code |= (u64)pre << 36;
    else if (meta & META_DYN) code |= (u64)pre;
    else code |= (u64)pre << (meta & META_LEN_MASK);      /* body length <= 64 - prefix_max */
    meta += (u32)pn;
}

/*
 * Merge step: consecutive plain codes of a thread are concatenated into one <= 64-bit item so
 * that the bit writer runs once per item instead of once per macroblock.  (acc, accn) is the
 * running item; `seal` tells the caller to store it (at the previous index) before this code
 * starts a new one.
 */
HSD bool merge_plain(u64 &acc, u32 &accn, u64 code, u32 n, u64 &sealed, u32 &sealed_n) {
    bool seal = accn + n > 64u;
    if (seal) { sealed = acc; sealed_n = accn; acc = 0; accn = 0; }
    acc = n < 64u ? ((acc << n) | code) : code;              /* n == 64 only ever follows a seal (acc == 0) */
    accn += n;
    return seal;
}

/* ---- sequential bit writer ------------------------------------------- */

struct BitPut {
    u32 acc;      /* bits of the word being assembled, from the MSB down */
    u32 fill;     /* bits used in acc, 0..31 */
    u32 j;        /* frame word index of acc */
    u32 head;     /* the string started mid-word and that first word is not complete yet */
    u32 H;        /* head partial: this thread's bits of the word its string starts in */
};

HSD void bp_init(BitPut &p, u32 bit) {
    p.acc = 0; p.fill = bit & 31u; p.j = bit >> 5; p.head = p.fill != 0u; p.H = 0;
}

HSD void bp_flush(BitPut &p, u32 word, u32 *win, u32 wbase, u32 wcap) {
    if (p.head) { p.H = word; p.head = 0; }
    else if (p.j - wbase < wcap) win[p.j - wbase] = word;
    p.j++;
}

/* append the low n (1..64) bits of code (bits above n must be zero) */
HSD void bp_put(BitPut &p, u64 code, u32 n, u32 *win, u32 wbase, u32 wcap) {
    const u64 t = code << (64u - n);
    const u64 a = ((u64)p.acc << 32) | (t >> p.fill);
    const u32 total = p.fill + n;
    if (total < 32u) { p.acc = (u32)(a >> 32); p.fill = total; return; }
    bp_flush(p, (u32)(a >> 32), win, wbase, wcap);
    if (total < 64u) { p.acc = (u32)a; p.fill = total - 32u; return; }
    bp_flush(p, (u32)a, win, wbase, wcap);
    p.acc = p.fill ? (u32)((t << (64u - p.fill)) >> 32) : 0u;
    p.fill = total - 64u;
}

/* ---- emulation prevention on 16-byte groups --------------------------- */

/* zero-byte flags of an MSB-first word as a nibble: bit k = k-th stream byte is zero */
HSD u32 zero_nib(u32 w) { return ((zero_byte_mask(w) >> 7) * 0x80402010u) >> 28; }

/*
 * Insert mask of a group (src/nal.c:24-50 in closed form): z = zero-byte mask (bit i = stream
 * byte i), m3 = bytes <= 3, vb = valid bytes (1..16), rin = zero bytes immediately before the
 * group.  0x03 goes before byte p iff byte p <= 3 and the zero run ending at p-1 has an even
 * length >= 2.  Runs are classified by the parity of their start with one carry-propagating add.
 */
HSD u32 epb_group_mask(u32 z, u32 m3, int vb, u32 rin) {
    const u32 r = rin >= 2u ? 2u + (rin & 1u) : rin;       /* keeps ">= 2" and the parity */
    const u32 valid = (1u << vb) - 1u;
    const u32 ze = ((z & valid) << 4) | (((1u << r) - 1u) << (4u - r));
    const u32 s = ze & ~(ze << 1);                          /* run starts */
    const u32 re = ze & ~(ze + (s & 0x55555555u));          /* bytes of runs that start at an even position */
    const u32 ro = ze & ~re;
    const u32 e = (((re << 1) & (re << 2)) & 0x55555555u) | (((ro << 1) & (ro << 2)) & 0xAAAAAAAAu);
    return (e >> 4) & m3 & valid;
}

/* zero bytes closing a group of vb valid bytes (rin added when the whole group is zero) */
HSD u32 epb_group_trail(u32 z, int vb, u32 rin) {
    const u32 t = ~z & ((1u << vb) - 1u);
    return t ? (u32)(vb - 32 + clz32(t)) : (u32)vb + rin;
}

/*
 * Zero bytes immediately before window byte b (the run that rbsp_to_ebsp's zero_count has seen,
 * src/nal.c:29-47): scans the window backwards; a run reaching the window start continues with
 * the state carried over from the bytes emitted before (zrun_in).
 */
HSD u32 zeros_before(const u32 *win, u32 b, u32 zrun_in) {
    u32 r = 0;
    int i = (int)(b >> 2);
    const u32 k = b & 3u;
    if (k) {
        u32 v = win[i] >> (8u * (4u - k));
        for (u32 t = 0; t < k; t++, v >>= 8) {
            if (v & 0xFFu) return r;
            r++;
        }
    }
    for (i--; i >= 0; i--) {
        const u32 v = win[i];
        if (v) return r + ((v & 0xFFu) ? 0u : (v & 0xFF00u) ? 1u : (v & 0xFF0000u) ? 2u : 3u);
        r += 4u;
    }
    return r + zrun_in;
}

/* byte-permute selector: output word = stream bytes s..s+3 of the MSB-first pair (a, b) */
HSD u32 realign_selector(u32 s) { return (u32)((0x5670670170120123ull >> (16u * s)) & 0xFFFFu); }

} /* namespace h264scroll */
#endif
