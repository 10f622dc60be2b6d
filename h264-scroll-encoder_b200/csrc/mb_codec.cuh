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
#define HSD_NOINLINE __host__ __device__ __noinline__
#else
#define HSD static inline
#define HSD_NOINLINE static
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


/* ======================================================================
 * Compose kernel v4: row-aligned segments of 8 macroblocks per thread, hints staged in shared
 * memory, 32-bit codes per macroblock, branch-free multi-word concatenation, and a work queue
 * that spreads everything that is not a short inline code (spliced macroblocks, long codes,
 * separately prefixed codes, the slice trailer) over the lanes of the CTA.
 * Pure functions again, shared with tests/emu.
 * ==================================================================== */

/*
 * Canonical hints.  The kernel compares neighbours through (mv word, low byte of the key word):
 *   - a spliced intra macroblock (H264GPU_MB_DYN_INTRA) must carry mv 0 and ref 0xFF
 *     ("available, never matches": src/h264_writer.c:636-640 publishes ref -1, mv 0),
 *   - positions outside the picture are filled in by the kernel with HINT_Y_UNAVAIL.
 * h264gpu_raster_regions and h264gpu_compose_host canonicalise; dense maps handed straight to
 * h264gpu_compose_batch go through h264gpu_canonicalize_hints once.  Inter refs >= 0xF0 are invalid.
 */
#define KEY_INTRA      0xFFu
#define KEY_UNAVAIL    0xFEu
#define HINT_Y_UNAVAIL 0x0000FFFEu        /* kind 0xFF, ref KEY_UNAVAIL, mv word 0 */

HSD u32 canonical_hint_x(u32 x, u32 y) { return ((y >> 8) & 0xFFu) == H264GPU_MB_DYN_INTRA ? 0u : x; }
HSD u32 canonical_hint_y(u32 y) { return ((y >> 8) & 0xFFu) == H264GPU_MB_DYN_INTRA ? (y | KEY_INTRA) : y; }

/* per-macroblock kinds held by the owning thread (4 bits each) */
#define MBK_NONE  0u      /* absent or P_Skip: no bits of its own */
#define MBK_SHORT 1u      /* whole code (with its mb_skip_run prefix) in <= 32 bits */
#define MBK_LONG  2u      /* inline code longer than that: code = packed predictor, rebuilt by the consumer */
#define MBK_DYN   3u      /* spliced: code = payload bit offset (len = prefix + payload bits) */
#define MBK_PFX   4u      /* short body whose mb_skip_run prefix does not fit beside it (SPEC skip only) */

/* queue items: x = slice bit | type << 30 */
#define IT_BITS 0u        /* y = prefix value (0: none), z = code, w = code bits */
#define IT_DYN  1u        /* y = prefix value, z = payload bit offset, w = payload bits */
#define IT_LONG 2u        /* y = prefix value | ref << 24, z = mvd_x, w = mvd_y */
#define IT_BIT_MASK 0x3FFFFFFFu

struct Item { u32 x, y, z, w; };

/* high word of (hi:lo) << n, 0 <= n <= 32 (n == 32 gives lo) */
HSD u32 fshl_c(u32 hi, u32 lo, u32 n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(lo, hi, n);
#else
    return n >= 32u ? lo : (n ? (hi << n) | (lo >> (32u - n)) : hi);
#endif
}

/* index of the leading one (v != 0) */
HSD u32 flo32(u32 v) { return 31u - (u32)clz32(v); }

/* se(v): value written (code number + 1) and its bit count (src/bitwriter.c:50-101) */
HSD u32 se_val(int v, u32 &n) {
    const u32 a = (u32)(v < 0 ? -v : v);
    const u32 cv = (a << 1) | (v <= 0 ? 1u : 0u);          /* v > 0: 2v, else 1 - 2v */
    n = 2u * flo32(cv) + 1u;
    return cv;
}

/* te(ref) of write_p16x16_mb (src/h264_writer.c:439-445): value and bit count */
HSD u32 ref_val(u32 ref, int num_refs, u32 &rl) {
    if (num_refs == 2) { rl = 1u; return ~ref & 1u; }
    if (num_refs < 2) { rl = 0u; return 0u; }
    const u32 rc = ref + 1u;
    rl = 2u * flo32(rc) + 1u;
    return rc;
}

/*
 * One inter macroblock against its neighbours A, B, C (C already replaced by D where the
 * reference does so; unavailable positions carry HINT_Y_UNAVAIL / mv 0).
 * Skip OFF: the code starts with the constant mb_skip_run bit (src/h264_writer.c:629-631).
 * Skip SPEC: body only, the prefix is attached once the run is known.
 */
HSD u32 decide_mb(u32 hx, u32 hy, u32 ax, u32 ay, u32 bx, u32 by, u32 cx, u32 cy,
                  int num_refs, int pred_spec, int skip_spec, u32 &code, u32 &len) {
    const u32 key = hy & 0xFFu;
    u32 ka = ay & 0xFFu, kb = by & 0xFFu, kc = cy & 0xFFu;
    const bool b_zero = kb == KEY_UNAVAIL || (kb == 0u && bx == 0u);   /* B as the P_Skip rule sees it */
    if (pred_spec && ka != KEY_UNAVAIL && kb == KEY_UNAVAIL && kc == KEY_UNAVAIL) { bx = ax; cx = ax; kb = ka; kc = ka; }
    const int ma = ka == key, mb = kb == key, mc = kc == key;
    u32 med = vmin2(vmax2(ax, bx), cx);                     /* the reference's median3 (:362-367) */
    if (pred_spec) med = vmax2(vmin2(ax, bx), med);
    const u32 single = ma ? ax : (mb ? bx : cx);
    const u32 pred = (ma + mb + mc == 1) ? single : med;
    if (skip_spec && key == 0u) {                           /* H.264 8.4.1.1 */
        const bool zero = ka == KEY_UNAVAIL || (ka == 0u && ax == 0u) || b_zero;
        if (hx == (zero ? 0u : pred)) { code = 0u; len = 0u; return MBK_NONE; }
    }
    const int dx = (int)(int16_t)(hx & 0xFFFFu) - (int)(int16_t)(pred & 0xFFFFu);
    const int dy = (int)(int16_t)(hx >> 16) - (int)(int16_t)(pred >> 16);
    u32 nx, ny, rl;
    const u32 vx = se_val(dx, nx), vy = se_val(dy, ny);
    const u32 rc = ref_val(key, num_refs, rl);
    const u32 tl = skip_spec ? 1u : 2u;
    len = tl + rl + nx + ny + 1u;
    if (len <= 32u) {
        u32 v = ((skip_spec ? 1u : 3u) << rl) | rc;
        v = (v << nx) | vx;
        v = (v << ny) | vy;
        code = (v << 1) | 1u;
        return MBK_SHORT;
    }
    code = pred;
    return MBK_LONG;
}

/* the same when A, B and C all equal the macroblock's own inter hint: mvd = 0, P_Skip with ref 0 */
HSD u32 decide_mb_uniform(u32 hy, int num_refs, int skip_spec, u32 &code, u32 &len) {
    const u32 key = hy & 0xFFu;
    if (skip_spec && key == 0u) { code = 0u; len = 0u; return MBK_NONE; }
    u32 rl;
    const u32 rc = ref_val(key, num_refs, rl);
    code = (((((skip_spec ? 1u : 3u) << rl) | rc) << 3) | 7u);
    len = (skip_spec ? 1u : 2u) + rl + 3u;
    return MBK_SHORT;
}

/* append the low n (0..32) bits of code to the right-aligned NWc-word integer a[] (a[NWc-1] least
   significant); step = how many codes went in before (words above step+1 are still zero) */
template <int NWc>
HSD void concat_append(u32 (&a)[NWc], u32 code, u32 n, int step) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NWc - 1; i++)
        if (i >= NWc - 1 - step) a[i] = fshl_c(a[i], a[i + 1], n);
    a[NWc - 1] = fshl_c(a[NWc - 1], 0u, n) | code;
}

/* b[0..NWc] = a[] shifted left by r (0..31): b[NWc] is the least significant word */
template <int NWc>
HSD void concat_place(const u32 (&a)[NWc], u32 r, u32 (&b)[NWc + 1]) {
    b[0] = fshl(0u, a[0], (int)r);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < NWc; i++) b[i] = fshl(a[i - 1], a[i], (int)r);
    b[NWc] = a[NWc - 1] << r;
}

#if defined(__CUDA_ARCH__)
#define WIN_OR(p, v) atomicOr((p), (v))
#else
#define WIN_OR(p, v) (*(p) |= (v))
#endif

/* OR the low n (1..64) bits of v into the window at slice bit `bit`; words outside [wbase, wbase+wcap) are dropped */
HSD void win_or64(u32 *win, u32 wbase, u32 wcap, u32 bit, u64 v, u32 n) {
    const u32 j = (bit >> 5) - wbase, s = bit & 31u;
    const u64 t = v << (64u - n);                           /* left aligned */
    const u64 u = t >> s;                                   /* bits for words j, j+1 */
    const u32 lo2 = s ? (u32)((t << (64u - s)) >> 32) : 0u; /* bits for word j+2 */
    if (j < wcap && (u32)(u >> 32)) WIN_OR(&win[j], (u32)(u >> 32));
    if (j + 1u < wcap && (u32)u) WIN_OR(&win[j + 1u], (u32)u);            /* u32 index: wraps like the check */
    if (j + 2u < wcap && lo2) WIN_OR(&win[j + 2u], lo2);
}

/* payload word wi as an MSB-first word (payload bytes are a bit stream: big-endian words) */
HSD u32 pay_word(const u32 *pw, u32 n_words, u32 wi) {
#if defined(__CUDA_ARCH__)
    return wi < n_words ? __byte_perm(pw[wi], 0u, 0x0123u) : 0u;      /* (generic load: the payload may sit in shared memory) */
#else
    return wi < n_words ? bswap32(pw[wi]) : 0u;
#endif
}
/* the same without the bounds check */
HSD u32 pay_raw(const u32 *p) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(*p, 0u, 0x0123u);
#else
    return bswap32(*p);
#endif
}
HSD u32 pay_ld32(const u32 *pw, u32 n_words, u32 p) {
    const u32 wi = p >> 5;
    return fshl(pay_word(pw, n_words, wi), pay_word(pw, n_words, wi + 1u), (int)(p & 31u));
}

/* copy `len` payload bits from payload bit `src` to slice bit `dst` (copy_bits, src/h264_writer.c:228-240):
   whole window words are plain stores, the partial words at both ends are OR-ed in */
HSD void splice_bits(u32 *win, u32 wbase, u32 wcap, u32 dst, const u32 *pw, u32 n_words, u32 src, u32 len) {
    const u32 s = dst & 31u;
    if (s && len) {
        const u32 hb = len < 32u - s ? len : 32u - s;
        const u32 v = (pay_ld32(pw, n_words, src) >> (32u - hb)) << (32u - s - hb);
        const u32 j = (dst >> 5) - wbase;
        if (j < wcap && v) WIN_OR(&win[j], v);
        dst += hb; src += hb; len -= hb;
    }
    u32 j = (dst >> 5) - wbase;
    const u32 nw = len >> 5, tb = len & 31u;
    if (nw) {
        u32 wi = src >> 5;
        const u32 sh = src & 31u;
        if (j < wcap && j + nw <= wcap && wi + nw < n_words) {
            /* the whole run lies inside the window and the payload: no per-word checks */
            const u32 *p = pw + wi;
            u32 *d = win + j;
            u32 a = pay_raw(p);
            u32 i = 0;
            for (; i + 4u <= nw; i += 4u) {                  /* four independent loads in flight */
                const u32 b0 = pay_raw(p + i + 1u), b1 = pay_raw(p + i + 2u), b2 = pay_raw(p + i + 3u), b3 = pay_raw(p + i + 4u);
                d[i] = fshl(a, b0, (int)sh);
                d[i + 1u] = fshl(b0, b1, (int)sh);
                d[i + 2u] = fshl(b1, b2, (int)sh);
                d[i + 3u] = fshl(b2, b3, (int)sh);
                a = b3;
            }
            for (; i < nw; i++) {
                const u32 b = pay_raw(p + i + 1u);
                d[i] = fshl(a, b, (int)sh);
                a = b;
            }
        } else {
            u32 a = pay_word(pw, n_words, wi);
            for (u32 i = 0; i < nw; i++) {
                const u32 b = pay_word(pw, n_words, ++wi);
                if (j + i < wcap) win[j + i] = fshl(a, b, (int)sh);
                a = b;
            }
        }
    }
    if (tb) {
        const u32 v = (pay_ld32(pw, n_words, src + 32u * nw) >> (32u - tb)) << (32u - tb);
        if (j + nw < wcap && v) WIN_OR(&win[j + nw], v);
    }
}

/* one queue item -> window */
HSD void consume_item(const Item &it, u32 *win, u32 wbase, u32 wcap, int num_refs, const u32 *pw, u32 n_words) {
    u32 bit = it.x & IT_BIT_MASK;
    const u32 type = it.x >> 30;
    u32 pre = type == IT_LONG ? (it.y & 0xFFFFFFu) : it.y;
    u32 src = it.z, n = it.w;
    const bool fused = type == IT_DYN && pre == 1u && src != 0u;
    const u32 bit0 = bit;
    if (fused) { src -= 1u; n += 1u; pre = 0u; }              /* mb_skip_run = 0, a single '1' in front: splice from one bit early, set that bit afterwards */
    if (pre) {
        const u32 pn = 2u * flo32(pre) + 1u;                 /* ue(run): run + 1 in 2m+1 bits */
        win_or64(win, wbase, wcap, bit, (u64)pre, pn);
        bit += pn;
    }
    if (type == IT_BITS) {
        if (n) win_or64(win, wbase, wcap, bit, (u64)src, n);
    } else if (type == IT_DYN) {
        splice_bits(win, wbase, wcap, bit, pw, n_words, src, n);
        const u32 j = (bit0 >> 5) - wbase;
        if (fused && j < wcap) WIN_OR(&win[j], 0x80000000u >> (bit0 & 31u));
    } else {
        MbEntry e;
        int bl = encode_p16x16_body((int)(it.y >> 24), (int)it.z, (int)it.w, num_refs, e);
        const u32 ws[3] = {e.w1, e.w2, e.w3};
        for (int i = 0; i < 3 && bl > 0; i++, bl -= 32) {
            const u32 n = (u32)(bl < 32 ? bl : 32);
            win_or64(win, wbase, wcap, bit, (u64)(ws[i] >> (32u - n)), n);
            bit += n;
        }
    }
}

/*
 * mb_skip_run prefixes for the coded macroblocks of a thread (SPEC skip): g0 = raster index of the
 * thread's first macroblock, prev = raster index of the last coded macroblock before it.  A short code
 * takes its prefix in when both fit 32 bits, else it becomes MBK_PFX (prefix travels separately).
 * Returns the raster index of the thread's last coded macroblock (or prev).
 */
HSD int attach_runs(u32 (&code)[8], u32 (&len)[8], u32 &kinds, int g0, int prev) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 8; k++) {
        const u32 kd = (kinds >> (4 * k)) & 0xFu;
        if (kd != MBK_NONE) {
            const u32 pre = (u32)(g0 + k - prev);           /* run + 1 */
            const u32 pn = 2u * flo32(pre) + 1u;
            if (kd == MBK_SHORT) {
                if (pn + len[k] <= 32u) code[k] |= pre << len[k];
                else kinds += (MBK_PFX - MBK_SHORT) << (4 * k);
            }
            len[k] += pn;
            prev = g0 + k;
        }
    }
    return prev;
}

/* true when a thread's macroblocks cannot all go through the concatenation path */
HSD bool kinds_complex(u32 kinds) { return (kinds & 0xEEEEEEEEu) != 0u; }

/*
 * The queue item of macroblock k of a complex thread.  bit = slice bit of the macroblock's first bit,
 * pre = its mb_skip_run prefix value (run + 1), (hx, hy) = its own hint (for MBK_LONG).
 */
HSD Item make_item(u32 kd, u32 code, u32 len, u32 bit, u32 pre, u32 hx, u32 hy) {
    Item it;
    if (kd == MBK_SHORT) { it.x = bit | (IT_BITS << 30); it.y = 0u; it.z = code; it.w = len; }
    else if (kd == MBK_PFX) { it.x = bit | (IT_BITS << 30); it.y = pre; it.z = code; it.w = len - (2u * flo32(pre) + 1u); }
    else if (kd == MBK_DYN) { it.x = bit | (IT_DYN << 30); it.y = pre; it.z = code; it.w = len - (2u * flo32(pre) + 1u); }
    else {
        it.x = bit | (IT_LONG << 30);
        it.y = pre | ((hy & 0xFFu) << 24);
        it.z = (u32)((int)(int16_t)(hx & 0xFFFFu) - (int)(int16_t)(code & 0xFFFFu));
        it.w = (u32)((int)(int16_t)(hx >> 16) - (int)(int16_t)(code >> 16));
    }
    return it;
}

/* 16-byte chunk c of a staged hint tile lives at chunk c ^ ((c >> 3) & 7): conflict-free 128-bit reads
   for threads that each own 64 consecutive bytes */
HSD u32 hint_swizzle(u32 c) { return c ^ ((c >> 3) & 7u); }

} /* namespace h264scroll */
#endif
