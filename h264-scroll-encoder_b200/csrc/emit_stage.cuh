/*
 * emit_stage.cuh - emulation prevention + store of a shared-memory window (rbsp_to_ebsp + the byte copy of
 * nal_write_unit, src/nal.c:24-84), shared by the compose kernel and the slice-rewrite kernel.
 *
 * The window holds slice bytes as MSB-first 32-bit words.  Smem must provide: u32 win[]; u32 gi[EPASS], gpre[EPASS];
 * unsigned short elist[EPASS], elist2[EPASS]; u32 n_list, n_list2; uint2 scan2[]; u32 zrun[2]; u32 cap; uint8_t *fout.
 * NTH = threads of the CTA (all of them call emit_window).
 */
#ifndef H264SCROLL_EMIT_STAGE_CUH
#define H264SCROLL_EMIT_STAGE_CUH

#include "mb_codec.cuh"
#include "kernel_util.cuh"

namespace h264scroll {

constexpr int EMIT_EG = 4;                /* 16-byte groups per thread per pass */
#ifndef EMIT_CLASSIFY_UNROLL
#define EMIT_CLASSIFY_UNROLL 2
#endif
#ifndef EMIT_STORE_UNROLL
#define EMIT_STORE_UNROLL 2
#endif
#ifndef EMIT_DIRECT
#define EMIT_DIRECT 1                     /* 1: a pass without insertions stores its groups straight from the window (see emit_window) */
#endif
constexpr int EMIT_CU = EMIT_CLASSIFY_UNROLL, EMIT_SU = EMIT_STORE_UNROLL;   /* (by two: fully unrolled loops cost more in code size than they save, 0.871 vs 0.862 ms) */

/* out of line: the backward scan is rare */
static __device__ __noinline__ u32 zeros_before_win(const u32 *win, u32 b, u32 zrun_in) { return zeros_before(win, b, zrun_in); }

/* the same in front of group q (byte 16 q), a whole group per step: inside a run of zero bytes - an all-zero macroblock
   is two or three groups long - this is the longest lane of the masks phase, and its steps are dependent loads */
static __device__ __noinline__ u32 zeros_before_group(const u32 *win, int q, u32 zrun_in) {
    u32 r = 0;
    for (int i = q - 1; i >= 0; i--) {
        const uint4 w = ((const uint4 *)win)[i];
        if (w.x | w.y | w.z | w.w) {
            const u32 v = w.w ? w.w : w.z ? w.z : w.y ? w.y : w.x;          /* the last word that is not zero */
            r += w.w ? 0u : w.z ? 4u : w.y ? 8u : 12u;
            return r + ((v & 0xFFu) ? 0u : (v & 0xFF00u) ? 1u : (v & 0xFF0000u) ? 2u : 3u);
        }
        r += 16u;
    }
    return r + zrun_in;
}

/* exact insert mask of group q of the window (src/nal.c:24-50 in closed form) */
__device__ __forceinline__ u32 group_insert_mask(const u32 *win, int q, uint4 w, int vb, u32 chunk_byte0, u32 zrun_in) {
    const u32 valid = (1u << vb) - 1u;
    u32 z = (zero_nib(w.x) | (zero_nib(w.y) << 4) | (zero_nib(w.z) << 8) | (zero_nib(w.w) << 12)) & valid;
    u32 keep = valid;
    if (chunk_byte0 == 0u && q == 0) { z &= ~0x1Fu; keep &= ~0x1Fu; }   /* start code + NAL header */
    const bool need = ((w.x >> 26) == 0u) && (keep & 1u);                /* a first byte <= 3 may close a zero run of the bytes before */
    if (!(need || (z & (z << 1)))) return 0u;
    u32 rin = 0;
    if (need && (q == 0 || (win[4 * q - 1] & 0xFFu) == 0u)) rin = zeros_before_group(win, q, zrun_in);
    const u32 m3 = zero_nib(w.x & 0xFCFCFCFCu) | (zero_nib(w.y & 0xFCFCFCFCu) << 4) |
                   (zero_nib(w.z & 0xFCFCFCFCu) << 8) | (zero_nib(w.w & 0xFCFCFCFCu) << 12);
    return epb_group_mask(z, m3 & keep, vb, rin);
}

/*
 * Classification test: can a byte INSIDE this group (MSB-first words w.x .. w.w = bytes 0..15) be an insertion point because
 * of two zero bytes that both lie in the group?  True when two neighbouring bytes of the group are zero (a superset: the byte
 * flags may also mark a 0x01 in front of a zero byte).  An insertion point behind zero bytes of the group BEFORE (byte 0 or 1)
 * shows as "first byte <= 3 behind a zero byte", which the callers test beside this.  A lone zero byte - one group in sixteen
 * of random data, most groups of a slice full of long Exp-Golomb codes - does not send the group to the masks phase.
 */
__device__ __forceinline__ bool group_zero_pair(uint4 w) {
    const u32 k = 0x01010101u, m = 0x80808080u;
    const u32 z0 = (w.x - k) & ~w.x & m, z1 = (w.y - k) & ~w.y & m, z2 = (w.z - k) & ~w.z & m, z3 = (w.w - k) & ~w.w & m;
    /* flag of byte j sits at bit 31 - 8 j of its word: a word's flags shifted up by one byte, the next word's first flag behind them */
    const u32 p = (z0 & __funnelshift_l(z1, z0, 8)) | (z1 & __funnelshift_l(z2, z1, 8)) |
                  (z2 & __funnelshift_l(z3, z2, 8)) | (z3 & (z3 << 8));
    return p != 0u;
}

constexpr u32 GI_CLEAN = 16u << 8;        /* sm.gi entry of a full group without insertions: ins << 16 | valid bytes << 8 | count */

/*
 * emit: emulation prevention + store of the window's first chunk_bytes bytes (frame bytes from
 * chunk_byte0 on).  Returns the bytes inserted; leaves the closing zero run in sm.zrun[zpar ^ 1].
 * A pass covers EG 16-byte groups per thread:
 *   classify  a group that holds two neighbouring zero bytes (or starts with a byte <= 3 behind a zero byte), the frame's
 *             last partial group, everything next to the slot's end -> list; all other groups are clean
 *   masks     one lane per listed group: exact insert mask + count
 *   offsets   prefix sum of the counts (packed 16-bit fields, one scan for the EG rounds)
 *   store     clean groups leave as aligned or re-aligned 32/128-bit words; the listed groups that
 *             do insert (or are partial) go byte by byte, again one lane per group
 * Called by all threads after a barrier; the caller puts one behind it.
 */
/* decoupled look-back over the chunks of one slice (slice rewrite): state word = flag << 30 | count; flag 1 = the
   chunk's own insertions, 2 = inclusive sum over the slice so far.  Returns the insertions before chunk c. */
__device__ __forceinline__ u32 chunk_lookback(volatile u32 *state, int c, u32 mine) {
    if (c == 0) { state[0] = (2u << 30) | mine; return 0u; }
    state[c] = (1u << 30) | mine;
    __threadfence();
    u32 sum = 0;
    for (int p = c - 1;; ) {
        const u32 v = state[p];
        if ((v >> 30) == 0u) continue;                       /* not published yet: the chunk is running (tickets are taken in order) */
        sum += v & 0x3FFFFFFFu;
        if ((v >> 30) == 2u || p == 0) break;
        p--;
    }
    state[c] = (2u << 30) | (sum + mine);
    return sum;
}

template <class Smem, int NTH>
__device__ __noinline__ u32 emit_window(Smem &sm, u32 chunk_byte0, int chunk_bytes, u32 ins_total, u32 zpar,
                                        volatile u32 *lookback = nullptr, int lookback_c = 0) {
    constexpr int NT = NTH, NW = NTH / 32, EG = EMIT_EG, EPASS = EMIT_EG * NTH;
    constexpr u32 FULL = 0xFFFFFFFFu;
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 zrun_in = sm.zrun[zpar];
    const u32 cap = sm.cap;
    uint8_t *fout = sm.fout;
    u32 *fout32 = (u32 *)fout;
    const u32 ins_in = ins_total;
    const int n_groups = (chunk_bytes + 15) >> 4;
    for (int q0 = 0; q0 < n_groups; q0 += EPASS) {
        /* even with an insertion after every second byte the pass stays clear of the slot's end */
        /* (slice rewrite: the insertions before the chunk are not known yet - at most one per two bytes) */
        const bool roomy = chunk_byte0 + 16u * (u32)q0 + 24u * (u32)min(EPASS, n_groups - q0) +
                           (lookback ? chunk_byte0 / 2u + 1u : ins_total) + 4u <= cap;
        /* ---- classify ---- */
        if (tid == 0) sm.n_list2 = 0;                        /* (read last in front of the previous pass's closing barrier) */
        /* rounds of NT groups this pass holds (a 720p scroll frame is a tenth of a window: one round, not four) */
        const int n_rounds = min(EG, (n_groups - q0 + NT - 1) / NT);
#pragma unroll EMIT_CU                         /* (code size matters more than the loop overhead) */
        for (int r = 0; r < EG; r++) {
            if (r >= n_rounds) break;
            const int gl = tid + NT * r, q = q0 + gl;
            const int vb = min(max(chunk_bytes - 16 * q, 0), 16);
            bool slow = false;
            if (vb > 0) {
                const uint4 w = ((const uint4 *)sm.win)[q];
                slow = group_zero_pair(w) || vb < 16 || !roomy;
                if (!slow && (w.x >> 26) == 0u) slow = q == 0 || (sm.win[4 * q - 1] & 0xFFu) == 0u;
            }
            sm.gi[gl] = (u32)vb << 8;
            const u32 m = __ballot_sync(FULL, slow);
            if (m) {
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&sm.n_list, (u32)__popc(m));
                base = __shfl_sync(FULL, base, 0);
                if (slow) sm.elist[base + (u32)__popc(m & ((1u << lane) - 1u))] = (unsigned short)gl;
            }
        }
        /* the zero run the window ends in goes to the next window (one thread, once: not a test in every round of the loop) */
        if (tid == NT - 1 && q0 + EPASS >= n_groups && chunk_bytes > 0) sm.zrun[zpar ^ 1u] = zeros_before_win(sm.win, (u32)chunk_bytes, zrun_in);
        TRACE(30);
        __syncthreads();
        TRACE(31);
        const u32 nl = sm.n_list;
        /* ---- masks ---- */
        for (u32 i = (u32)tid; i < nl; i += NT) {
            const int gl = sm.elist[i], q = q0 + gl;
            const int vb = min(chunk_bytes - 16 * q, 16);
            const u32 ins = group_insert_mask(sm.win, q, ((const uint4 *)sm.win)[q], vb, chunk_byte0, zrun_in);
            const u32 g = (ins << 16) | ((u32)vb << 8) | (u32)__popc(ins);
            sm.gi[gl] = g;
            /* most listed groups turn out clean (a lone zero byte): the few that do insert - or are partial, or sit next
               to the slot's end - go into a second list, so that the byte-wise phase is as short as its real work */
            /* (counter: entries in the low half; the high half counts the groups that really insert or sit next to the
               slot's end - a pass without any takes the direct way out below) */
            if (!(roomy && g == GI_CLEAN)) sm.elist2[atomicAdd(&sm.n_list2, (ins != 0u || !roomy) ? 0x10001u : 1u) & 0xFFFFu] = (unsigned short)gl;
        }
        TRACE(32);
        __syncthreads();
        TRACE(33);
        if (tid == 0) sm.n_list = 0;                         /* the next pass adds to it only after the barriers below */
        const u32 l2 = sm.n_list2;
#if EMIT_DIRECT
        /* ---- nothing is inserted in this pass and the bytes before it sit on a 16-byte boundary of the slot: every group
                goes out where it lies, as one 128-bit store - no offsets, no lists (the common case: an insertion needs
                00 00 0x in the slice, and a frame's only window starts at byte 0 of its slot) ---- */
        if ((l2 >> 16) == 0u && lookback == nullptr && (ins_total & 15u) == 0u) {
            uint4 *dst4 = (uint4 *)(fout + chunk_byte0 + ins_total);
#pragma unroll EMIT_SU
            for (int r = 0; r < EG; r++) {
                if (r >= n_rounds) break;
                const int q = q0 + tid + NT * r;
                if (16 * (q + 1) <= chunk_bytes) {
                    const uint4 w = ((const uint4 *)sm.win)[q];
                    dst4[q] = make_uint4(__byte_perm(w.x, 0u, 0x0123u), __byte_perm(w.y, 0u, 0x0123u),
                                         __byte_perm(w.z, 0u, 0x0123u), __byte_perm(w.w, 0u, 0x0123u));
                }
            }
            const int ql = chunk_bytes >> 4;                 /* the chunk's partial last group, byte by byte */
            if (tid < (chunk_bytes & 15) && ql >= q0 && ql < q0 + EPASS)
                fout[chunk_byte0 + ins_total + 16u * (u32)ql + (u32)tid] = (uint8_t)(sm.win[4 * ql + (tid >> 2)] >> (24 - 8 * (tid & 3)));
            if (q0 + EPASS < n_groups) __syncthreads();
            continue;
        }
#endif
        /* ---- offsets ---- */
        u32 lo = sm.gi[tid] & 0xFFu, hi = 0u;                /* (rounds past n_rounds were not classified: nothing there) */
        if (n_rounds > 1) lo |= (sm.gi[tid + NT] & 0xFFu) << 16;
        if (n_rounds > 2) hi = sm.gi[tid + 2 * NT] & 0xFFu;
        if (n_rounds > 3) hi |= (sm.gi[tid + 3 * NT] & 0xFFu) << 16;
        u32 tlo, thi;
        group_excl_scan_x2<NW, 0>(lo, hi, sm.scan2, tid >> 5, tlo, thi);      /* barriers on both sides of every pass: one buffer does */
        TRACE(34);
        const u32 t0 = tlo & 0xFFFFu, t1 = tlo >> 16, t2 = thi & 0xFFFFu, t3 = thi >> 16;
        if (lookback) {
            /* slice rewrite: where this chunk's bytes go depends on the insertions of the chunks before it (single pass
               per chunk: EPASS groups).  One thread looks back, everybody picks the result up from shared memory */
            if (tid == 0) sm.scan2[NW].x = chunk_lookback(lookback, lookback_c, t0 + t1 + t2 + t3);
            __syncthreads();
            ins_total = sm.scan2[NW].x;
        }
        /* ---- store: clean groups ---- */
        u32 round_base = ins_total;                          /* bytes inserted before the round's first group */
#pragma unroll EMIT_SU
        for (int r = 0; r < EG; r++) {
            if (r >= n_rounds) break;
            const int gl = tid + NT * r, q = q0 + gl;
            const u32 g = sm.gi[gl];
            const u32 before = round_base + (((r & 2) ? hi : lo) >> (16 * (r & 1)) & 0xFFFFu);
            round_base += ((r & 2) ? thi : tlo) >> (16 * (r & 1)) & 0xFFFFu;
            if ((g >> 8) == 0u) continue;                    /* past the chunk */
            if (!(roomy && g == GI_CLEAN)) { sm.gpre[gl] = before; continue; }
            const u32 D = chunk_byte0 + 16u * (u32)q + before;         /* output byte of the group's first byte */
            const uint4 w = ((const uint4 *)sm.win)[q];
            const u32 s = (4u - (D & 3u)) & 3u;              /* source bytes before the first aligned word */
            u32 *dst = fout32 + ((D + s) >> 2);
            if (s == 0u) {
                const uint4 o = make_uint4(__byte_perm(w.x, 0u, 0x0123u), __byte_perm(w.y, 0u, 0x0123u),
                                           __byte_perm(w.z, 0u, 0x0123u), __byte_perm(w.w, 0u, 0x0123u));
                if ((D & 15u) == 0u) *(uint4 *)dst = o;
                else { dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w; }
            } else {
                /* re-aligned words; the bytes shared with a neighbour are whole-word stores only when the neighbour is clean too */
                const bool prev_clean = gl > 0 && sm.gi[gl - 1] == GI_CLEAN;
                const bool next_clean = gl < EPASS - 1 && q + 1 < n_groups && sm.gi[gl + 1] == GI_CLEAN;   /* (classified in this pass) */
                const u32 wn = sm.win[4 * q + 4];
                const u32 sel = realign_selector(s);
                dst[0] = __byte_perm(w.x, w.y, sel);
                dst[1] = __byte_perm(w.y, w.z, sel);
                dst[2] = __byte_perm(w.z, w.w, sel);
                const u32 o3 = __byte_perm(w.w, wn, sel);
                if (next_clean) dst[3] = o3;
                else {                                       /* (1..3 bytes, no loop) */
                    uint8_t *t = fout + D + s + 12u;
                    t[0] = (uint8_t)o3;
                    if (s < 3u) t[1] = (uint8_t)(o3 >> 8);
                    if (s < 2u) t[2] = (uint8_t)(o3 >> 16);
                }
                if (!prev_clean) {
                    uint8_t *h = fout + D;
                    h[0] = (uint8_t)(w.x >> 24);
                    if (s > 1u) h[1] = (uint8_t)(w.x >> 16);
                    if (s > 2u) h[2] = (uint8_t)(w.x >> 8);
                }
            }
        }
        TRACE(35);
        __syncthreads();
        TRACE(36);
        /* ---- store: listed groups that insert, are partial, or sit next to the slot's end ---- */
        const u32 nl2 = l2 & 0xFFFFu;
        for (u32 i = (u32)tid; i < 4u * nl2; i += NT) {       /* four lanes per group, a word each: the phase is as long as its longest lane */
            const int gl = sm.elist2[i >> 2], q = q0 + gl;
            const u32 sub = i & 3u;
            const u32 gg = sm.gi[gl];
            const u32 ins = gg >> 16;
            const int vbs = (int)((gg >> 8) & 0xFFu) - 4 * (int)sub;           /* valid bytes from this word on */
            if (vbs <= 0) continue;
            const u32 wv = sm.win[4 * q + (int)sub];
            u32 pos = chunk_byte0 + 16u * (u32)q + sm.gpre[gl] + 4u * sub + (u32)__popc(ins & ((1u << (4u * sub)) - 1u));
            u32 is = ins >> (4u * sub);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < vbs) {
                    if (is & 1u) { if (pos < cap) fout[pos] = 3; pos++; }
                    if (pos < cap) fout[pos] = (uint8_t)(wv >> (24 - 8 * k));
                    pos++;
                    is >>= 1;
                }
            }
        }
        TRACE(37);
        ins_total += t0 + t1 + t2 + t3;
        if (q0 + EPASS < n_groups) __syncthreads();          /* list, counts and offsets are reused by the next pass (the caller
                                                                 puts a barrier behind the last one) */
    }
    return ins_total - ins_in;
}

/*
 * The same stage, warp by warp: warp w owns groups [128 w, 128 w + 128) of the window (lane l those with index = l mod 32,
 * conflict-free 128-bit loads).  Flags and counts travel by ballot and shuffle instead of through shared memory, the rare
 * groups are listed per warp (positions from the ballots, no atomics) and worked through by the warp that owns them, and
 * the CTA meets ONCE, for the warps' insert totals - two barriers per window (with the caller's) instead of five.
 * One pass: the window must not hold more than 4 * NTH groups; a window that could touch the slot's end goes the careful way
 * (emit_window) - emit_window_fits tells (ins_bound = what may have been inserted before the window), the caller picks.
 * Used by the slice-rewrite kernel (0.197 -> see DESIGN.md 3.4); in the compose kernel it costs more in registers around
 * the call than it saves (profiles/r2_kernel_experiments.txt, w1-w3).
 */
template <class Smem, int NTH>
__device__ __forceinline__ bool emit_window_fits(const Smem &sm, u32 chunk_byte0, int chunk_bytes, u32 ins_bound) {
    const int n_groups = (chunk_bytes + 15) >> 4;
    return n_groups <= EMIT_EG * NTH && chunk_byte0 + 24u * (u32)n_groups + ins_bound + 4u <= sm.cap;
}
#ifndef EMIT_WARPS_INLINE
#define EMIT_WARPS_INLINE __noinline__
#endif
template <class Smem, int NTH>
__device__ EMIT_WARPS_INLINE u32 emit_window_warps(Smem &sm, u32 chunk_byte0, int chunk_bytes, u32 ins_total, u32 zpar,
                                                   volatile u32 *lookback = nullptr, int lookback_c = 0) {
    constexpr int NW = NTH / 32, EG = EMIT_EG;
    constexpr u32 FULL = 0xFFFFFFFFu;
    static_assert(EG == 4, "four rounds of 32 groups per warp");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lt = (1u << lane) - 1u;
    const u32 zrun_in = sm.zrun[zpar];
    uint8_t *fout = sm.fout;
    u32 *fout32 = (u32 *)fout;
    const int g0 = 128 * warp;                                   /* the warp's first group */
    unsigned short *wlist = sm.elist + g0, *wlist2 = sm.elist2 + g0;

    /* (the loops over the four rounds stay rolled, their results in shared memory: what this function keeps in registers
       the kernel around it has to give up across the call) */
    /* ---- classify: which groups need a look (a zero byte inside, a first byte <= 3 behind one, the partial last group) ---- */
    int n_slow = 0;
#pragma unroll 1
    for (int r = 0; r < EG; r++) {
        const int q = g0 + 32 * r + lane;
        const int vb = min(max(chunk_bytes - 16 * q, 0), 16);
        bool slow = false;
        if (vb > 0) {
            const uint4 w = ((const uint4 *)sm.win)[q];
            slow = group_zero_pair(w) || vb < 16;
            if (!slow && (w.x >> 26) == 0u) slow = q == 0 || (sm.win[4 * q - 1] & 0xFFu) == 0u;
            /* the holder of the window's last byte hands the zero run to the next window */
            if (16 * (q + 1) >= chunk_bytes) sm.zrun[zpar ^ 1u] = zeros_before_win(sm.win, (u32)chunk_bytes, zrun_in);
        }
        sm.gi[q] = (u32)vb << 8;
        const u32 m = __ballot_sync(FULL, slow);
        if (slow) wlist[n_slow + __popc(m & lt)] = (unsigned short)(32 * r + lane);
        n_slow += __popc(m);
    }
    __syncwarp();
    /* ---- masks: one lane per listed group of the warp ---- */
    for (int i = lane; i < n_slow; i += 32) {
        const int q = g0 + wlist[i];
        const int vb = min(chunk_bytes - 16 * q, 16);
        const u32 ins = group_insert_mask(sm.win, q, ((const uint4 *)sm.win)[q], vb, chunk_byte0, zrun_in);
        sm.gi[q] = (ins << 16) | ((u32)vb << 8) | (u32)__popc(ins);
    }
    __syncwarp();
    /* ---- counts: per lane four groups, packed 16-bit fields; scan over the lanes; the warp's total to the CTA ---- */
    u32 il = (sm.gi[g0 + lane] & 0xFFu) | ((sm.gi[g0 + 32 + lane] & 0xFFu) << 16);
    u32 ih = (sm.gi[g0 + 64 + lane] & 0xFFu) | ((sm.gi[g0 + 96 + lane] & 0xFFu) << 16);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 a = __shfl_up_sync(FULL, il, d), b = __shfl_up_sync(FULL, ih, d);
        if (lane >= d) { il += a; ih += b; }
    }
    const u32 tl = __shfl_sync(FULL, il, 31), th = __shfl_sync(FULL, ih, 31);      /* the rounds' totals */
    if (lane == 0) sm.scan2[warp].x = (tl & 0xFFFFu) + (tl >> 16) + (th & 0xFFFFu) + (th >> 16);
    __syncthreads();
    u32 v = lane < NW ? sm.scan2[lane].x : 0u;
    u32 round_base = lane < warp ? v : 0u;
#pragma unroll
    for (int d = 16; d; d >>= 1) { v += __shfl_xor_sync(FULL, v, d); round_base += __shfl_xor_sync(FULL, round_base, d); }
    const u32 total = v;
    if (lookback) {
        /* slice rewrite: the insertions of the chunks before this one - one thread looks back, everybody picks the result up */
        if (tid == 0) sm.scan2[NW].x = chunk_lookback(lookback, lookback_c, total);
        __syncthreads();
        ins_total = sm.scan2[NW].x;
    }
    round_base += ins_total;                                     /* bytes inserted before the round's first group */
    /* ---- store ---- */
    u32 n_ins = 0;
#pragma unroll 1
    for (int r = 0; r < EG; r++) {
        const int q = g0 + 32 * r + lane;
        const u32 g = sm.gi[q];
        const u32 incl = ((r & 2) ? ih : il) >> (16 * (r & 1)) & 0xFFFFu;
        const u32 before = round_base + incl - (g & 0xFFu);
        round_base += ((r & 2) ? th : tl) >> (16 * (r & 1)) & 0xFFFFu;
        const bool bytewise = (g >> 8) != 0u && g != GI_CLEAN;
        const u32 bm = __ballot_sync(FULL, bytewise);
        if (bytewise) {
            sm.gpre[q] = before;
            wlist2[n_ins + (u32)__popc(bm & lt)] = (unsigned short)(32 * r + lane);
        }
        n_ins += (u32)__popc(bm);
        if (g != GI_CLEAN) continue;
        const u32 D = chunk_byte0 + 16u * (u32)q + before;         /* output byte of the group's first byte */
        const uint4 w = ((const uint4 *)sm.win)[q];
        const u32 s = (4u - (D & 3u)) & 3u;                      /* source bytes before the first aligned word */
        u32 *dst = fout32 + ((D + s) >> 2);
        if (s == 0u) {
            const uint4 o = make_uint4(__byte_perm(w.x, 0u, 0x0123u), __byte_perm(w.y, 0u, 0x0123u),
                                       __byte_perm(w.z, 0u, 0x0123u), __byte_perm(w.w, 0u, 0x0123u));
            if ((D & 15u) == 0u) *(uint4 *)dst = o;
            else { dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w; }
        } else {
            /* re-aligned words; the bytes shared with a neighbour are whole-word stores only when the neighbour is clean too
               (the warp's first and last group do not look at the other warps' groups: bytes) */
            const int ql = 32 * r + lane;
            const bool prev_clean = ql > 0 && sm.gi[q - 1] == GI_CLEAN;
            const bool next_clean = ql < 127 && sm.gi[q + 1] == GI_CLEAN;
            const u32 wn = sm.win[4 * q + 4];
            const u32 sel = realign_selector(s);
            dst[0] = __byte_perm(w.x, w.y, sel);
            dst[1] = __byte_perm(w.y, w.z, sel);
            dst[2] = __byte_perm(w.z, w.w, sel);
            const u32 o3 = __byte_perm(w.w, wn, sel);
            if (next_clean) dst[3] = o3;
            else {                                               /* (1..3 bytes, no loop) */
                uint8_t *t = fout + D + s + 12u;
                t[0] = (uint8_t)o3;
                if (s < 3u) t[1] = (uint8_t)(o3 >> 8);
                if (s < 2u) t[2] = (uint8_t)(o3 >> 16);
            }
            if (!prev_clean) {
                uint8_t *h = fout + D;
                h[0] = (uint8_t)(w.x >> 24);
                if (s > 1u) h[1] = (uint8_t)(w.x >> 16);
                if (s > 2u) h[2] = (uint8_t)(w.x >> 8);
            }
        }
    }
    __syncwarp();
    /* ---- store: the groups that insert or are partial, four lanes each, a word per lane ---- */
    for (u32 i = (u32)lane; i < 4u * n_ins; i += 32u) {
        const int q = g0 + wlist2[i >> 2];
        const u32 sub = i & 3u;
        const u32 gg = sm.gi[q];
        const u32 ins = gg >> 16;
        const int vbs = (int)((gg >> 8) & 0xFFu) - 4 * (int)sub;               /* valid bytes from this word on */
        if (vbs <= 0) continue;
        const u32 wv = sm.win[4 * q + (int)sub];
        u32 pos = chunk_byte0 + 16u * (u32)q + sm.gpre[q] + 4u * sub + (u32)__popc(ins & ((1u << (4u * sub)) - 1u));
        u32 is = ins >> (4u * sub);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < vbs) {
                if (is & 1u) fout[pos++] = 3;
                fout[pos++] = (uint8_t)(wv >> (24 - 8 * k));
                is >>= 1;
            }
        }
    }
    return lookback ? ins_total + total : total;
}

}  /* namespace h264scroll */
#endif
