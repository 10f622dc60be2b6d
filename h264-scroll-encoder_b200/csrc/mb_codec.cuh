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

/* a long macroblock body, left aligned in three words (the low 11 bits of w3 are never body bits) */
struct alignas(16) MbEntry { u32 w0, w1, w2, w3; };
#define ENT_META_MASK 0x7FFu

/* append the low n (1..33) bits of v to the right-aligned accumulator hi:lo */
HSD void acc_push(u64 &hi, u64 &lo, u64 v, int n) {
    hi = (hi << n) | (lo >> (64 - n));
    lo = (lo << n) | v;
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

/* ---- emulation prevention on 16-byte groups ---------------------------
 *
 * rbsp_to_ebsp (src/nal.c:24-50) in closed form: with r = number of zero bytes immediately before RBSP byte p
 * (continuing the automaton state carried into the window), 0x03 is inserted before p iff byte(p) <= 3, r >= 2, r even.
 * Windows hold the stream MSB-first (stream byte 4j+k is bits [31-8k:24-8k] of word j); stream bytes below 5 are the
 * start code and NAL header and never take part.
 */

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

/* word wi of a bit source as an MSB-first word; sel = 0x0123 for payload bytes (a byte stream: big-endian words),
   0x3210 for words that are MSB-first already; indices outside [0, n_words) read as zero */
HSD u32 src_raw(const u32 *p, u32 sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(*p, 0u, sel);                         /* (generic load: the source may sit in shared memory) */
#else
    return sel == 0x0123u ? bswap32(*p) : *p;
#endif
}
HSD u32 src_word(const u32 *pw, u32 n_words, int wi, u32 sel) {
    return (u32)wi < n_words ? src_raw(pw + wi, sel) : 0u;
}

/*
 * copy `len` bits from bit `src` (>= -1: bit -1 reads as zero) of the MSB-first source pw[] to slice bit `dst`
 * (copy_bits, src/h264_writer.c:228-240), word by word of the DESTINATION: window word j takes
 * fshl(source word, next source word, shift) - one load, one funnel shift, one store per word, the same for every word;
 * the first and the last word are masked and OR-ed in (their other bits belong to the neighbours), first_or is added
 * to the first word.
 */
HSD void splice_bits(u32 *win, u32 wbase, u32 wcap, u32 dst, const u32 *pw, u32 n_words, int src, u32 len, u32 first_or, u32 sel) {
    if (!len) return;
    const u32 s = dst & 31u;
    const u32 nd = (s + len + 31u) >> 5;                     /* window words touched */
    const u32 j = (dst >> 5) - wbase;
    const int d0 = src - (int)s;                             /* source bit under bit 0 of the first window word (>= -32) */
    const int wi = d0 >> 5;                                  /* (arithmetic shift: floor) */
    const int sh = d0 & 31;
    const u32 e = (s + len) & 31u;
    const u32 mf = 0xFFFFFFFFu >> s, ml = e ? ~(0xFFFFFFFFu >> e) : 0xFFFFFFFFu;
    if (j < wcap && j + nd <= wcap && (u32)(wi + (int)nd) < n_words) {
        /* everything lies inside the window and the source: no per-word checks */
        const u32 *p = pw + wi;
        u32 *d = win + j;
        u32 a = wi >= 0 ? src_raw(p, sel) : 0u;
        u32 b = src_raw(p + 1, sel);
        u32 v = (fshl(a, b, sh) & mf) | first_or;
        if (nd == 1u) { WIN_OR(d, v & (ml | first_or)); return; }
        WIN_OR(d, v);
        a = b;
        u32 i = 1;
        for (; i + 4u < nd; i += 4u) {                       /* four independent loads in flight */
            const u32 b0 = src_raw(p + i + 1u, sel), b1 = src_raw(p + i + 2u, sel), b2 = src_raw(p + i + 3u, sel), b3 = src_raw(p + i + 4u, sel);
            d[i] = fshl(a, b0, sh);
            d[i + 1u] = fshl(b0, b1, sh);
            d[i + 2u] = fshl(b1, b2, sh);
            d[i + 3u] = fshl(b2, b3, sh);
            a = b3;
        }
        for (; i + 1u < nd; i++) {
            b = src_raw(p + i + 1u, sel);
            d[i] = fshl(a, b, sh);
            a = b;
        }
        v = fshl(a, src_raw(p + nd, sel), sh) & ml;
        if (v) WIN_OR(d + (nd - 1u), v);
    } else {
        u32 a = src_word(pw, n_words, wi, sel);
        for (u32 i = 0; i < nd; i++) {
            const u32 b = src_word(pw, n_words, wi + 1 + (int)i, sel);
            u32 v = fshl(a, b, sh);
            if (i == 0u) v = (v & mf) | first_or;
            if (i == nd - 1u) v &= ml | (i == 0u ? first_or : 0u);
            if (j + i < wcap && v) WIN_OR(&win[j + i], v);
            a = b;
        }
    }
}

/* one queue item -> window; self = the item's four words in memory (+ two readable words behind them) */
HSD void consume_item(const Item &it, const u32 *self, u32 *win, u32 wbase, u32 wcap, int num_refs, const u32 *pw, u32 n_words) {
    u32 bit = it.x & IT_BIT_MASK;
    const u32 type = it.x >> 30;
    u32 pre = type == IT_LONG ? (it.y & 0xFFFFFFu) : it.y;
    u32 first_or = 0u;
    int src = (int)it.z;
    u32 n = it.w;
    if (type == IT_DYN && pre == 1u) {
        /* mb_skip_run = 0, a single '1' in front: splice from one bit early, that bit set in the first word */
        first_or = 0x80000000u >> (bit & 31u);
        src -= 1; n += 1u; pre = 0u;
    }
    if (pre) {
        const u32 pn = 2u * flo32(pre) + 1u;                 /* ue(run): run + 1 in 2m+1 bits */
        win_or64(win, wbase, wcap, bit, (u64)pre, pn);
        bit += pn;
    }
    if (type != IT_LONG) {
        /* a short code is a one-word source: the item's own z word, its bits right-aligned */
        const bool dyn = type == IT_DYN;
        splice_bits(win, wbase, wcap, bit, dyn ? pw : self + 2, dyn ? n_words : 3u, dyn ? src : 32 - (int)n, n, first_or, dyn ? 0x0123u : 0x3210u);
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
