/*
 * compose_kernels.cu - sm_100a kernels of the P-slice composer.
 *
 * compose_frames_kernel: one CTA composes one frame at a time (grid-stride over
 * frames); a frame is walked in tiles of TILE = 256 threads x 8 macroblocks.
 * Everything per-macroblock lives in registers; shared memory only holds the
 * slice bits of the tile (the "window") on their way out.
 *
 *   load     every thread reads its 8 consecutive hints (4 x 128-bit loads) plus
 *            the 10 hints above and the one to the left straight from global memory
 *            (the rows above come back from L1/L2); nothing is staged
 *   decide   per MB: A/B/C(D) neighbour views, packed int16x2 predictor, P_Skip test,
 *            mvd, closed-form Exp-Golomb -> a <= 64-bit code + length in registers
 *            (replaces src/h264_writer.c:595-641, :369-453, src/bitwriter.c:50-101)
 *   skiprun  SPEC mode only: block prefix-max of "last coded MB" -> mb_skip_run prefixes
 *   scan     block prefix sum of the threads' bit counts -> slice bit offsets
 *   write    each thread appends its codes with a sequential bit writer: whole words
 *            are plain stores; the partial first/last words of neighbouring threads are
 *            merged with a segmented warp-shuffle OR, so shared-memory atomics are left
 *            for the two words a warp shares with its neighbours
 *   splice   whole words inside a spliced macroblock are funnel-shifted in from the
 *            pre-encoded payload by groups of 8 lanes (semantics of copy_bits,
 *            src/h264_writer.c:228); the partial words at both ends go through the
 *            owner thread's bit writer
 *   emit     emulation prevention (src/nal.c:24-50) on 16-byte groups: zero-byte masks,
 *            a bit-parallel run-parity rule, prefix sum of the insert counts; groups
 *            without insertions leave as re-aligned 32-bit words (one byte-permute each)
 *
 * The Annex-B start code, the NAL header byte and the host-written slice header are
 * the first bits of the same bit string, so the window image is the output image.
 */
#include <cuda_runtime.h>
#include "mb_codec.cuh"
#include "compose_kernels.cuh"

namespace h264scroll {

constexpr int NT = COMPOSE_THREADS;      /* threads per CTA */
constexpr int MPT = COMPOSE_MPT;         /* consecutive macroblocks per thread */
constexpr int TILE = NT * MPT;           /* macroblocks per tile */
constexpr int NW = NT / 32;
constexpr int CH = COMPOSE_WINDOW_WORDS; /* window: slice words staged per pass */
constexpr int LCAP = COMPOSE_COPY_SLOTS; /* splice copy list entries per pass */
constexpr u32 FULL = 0xFFFFFFFFu;
#ifndef COMPOSE_ABLATE
#define COMPOSE_ABLATE 0                 /* timing experiments only (scripts/variant.sh); 0 in the shipped build */
#endif
#ifndef COMPOSE_TIMING
#define COMPOSE_TIMING 0                 /* 1: thread 0 of every CTA accumulates clock64() per pipeline stage (experiments only) */
#endif
#if COMPOSE_TIMING
__device__ unsigned long long g_stage_cycles[8];
#define STAGE_MARK(i) do { if (tid == 0) { const long long t_ = clock64(); stage_acc[i] += t_ - stage_t; stage_t = t_; } } while (0)
#else
#define STAGE_MARK(i) do { } while (0)
#endif

struct __align__(16) ComposeSmem {
    u32 win[CH + 8];                     /* slice words [wbase, wbase + CH) of the frame, MSB first (+ read-ahead pad); zero where unwritten */
    uint4 copy[LCAP];                    /* {first window word, payload bit, words, -} */
    u32 gi[NT];                          /* emit: (insert count << 8) | valid bytes of each group */
    u32 scan[2][NW + 1];                 /* double buffered: one barrier per block scan */
    u32 zrun[2];                         /* zero bytes closing the bytes emitted so far (rbsp_to_ebsp state), double buffered */
    u32 n_copy[2];                       /* queued splice runs, double buffered per tile */
    /* per-frame constants only the splice and emit stages need: kept out of the register file */
    const u32 *tab;                      /* the frame's payload bit-offset table */
    const u32 *pay_words;                /* the frame's payload */
    u32 n_pay_words;
    u32 cap;                             /* output slot capacity */
    uint8_t *fout;                       /* output slot */
};

/*
 * Exclusive prefix sum over the CTA; every thread gets its offset and the total.  One barrier:
 * consecutive scans must alternate between the two s_scan buffers.
 */
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *s_scan, u32 &total, int pred, int &any_pred) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_scan[warp] = inc;
    any_pred = __syncthreads_or(pred);                 /* the scan's barrier doubles as a block-wide vote */
    const u32 wsum = lane < NW ? s_scan[lane] : 0u;    /* every warp scans the NW warp totals itself */
    u32 winc = wsum;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, winc, d);
        if (lane >= d) winc += t;
    }
    const u32 wexc = __shfl_sync(FULL, winc - wsum, warp);
    total = __shfl_sync(FULL, winc, NW - 1);
    return wexc + inc - v;
}

/* exclusive prefix max over the CTA (identity -1 as int); `total` = max over all threads */
__device__ __noinline__ int block_excl_scan_max(int v, int *s_scan, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = max(inc, t);
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    const int wv = lane < NW ? s_scan[lane] : -1;
    int winc = wv;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        int t = __shfl_up_sync(FULL, winc, d);
        if (lane >= d) winc = max(winc, t);
    }
    int wexc = __shfl_up_sync(FULL, winc, 1);
    if (lane == 0) wexc = -1;
    wexc = __shfl_sync(FULL, wexc, warp);
    total = __shfl_sync(FULL, winc, NW - 1);
    int exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = -1;
    return max(wexc, exc);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* payload word wi as an MSB-first word (payload bytes are a bit stream, so big-endian in memory) */
__device__ __forceinline__ u32 payload_word(const u32 *__restrict__ pw, u32 n_words, u32 wi) {
    return wi < n_words ? __byte_perm(__ldg(pw + wi), 0u, 0x0123u) : 0u;
}

/* 32 payload bits from bit position p */
__device__ __forceinline__ u32 payload_ld32(const u32 *__restrict__ pw, u32 n_words, u32 p) {
    const u32 wi = p >> 5;
    return __funnelshift_l(payload_word(pw, n_words, wi + 1u), payload_word(pw, n_words, wi), p & 31u);
}

/* n whole slice words of a spliced macroblock: window word j.. <- payload bits p.. (copy_bits, src/h264_writer.c:228) */
__device__ __forceinline__ void splice_words(u32 *win, u32 j, u32 n, const u32 *__restrict__ pw, u32 n_words, u32 p) {
    u32 wi = p >> 5;
    const u32 sh = p & 31u;
    u32 a = payload_word(pw, n_words, wi);
    u32 i = 0;
    for (; i + 4u <= n; i += 4u) {                          /* four independent loads in flight */
        const u32 b0 = payload_word(pw, n_words, wi + 1u), b1 = payload_word(pw, n_words, wi + 2u),
                  b2 = payload_word(pw, n_words, wi + 3u), b3 = payload_word(pw, n_words, wi + 4u);
        if (j + i < (u32)CH) win[j + i] = __funnelshift_l(b0, a, sh);
        if (j + i + 1u < (u32)CH) win[j + i + 1u] = __funnelshift_l(b1, b0, sh);
        if (j + i + 2u < (u32)CH) win[j + i + 2u] = __funnelshift_l(b2, b1, sh);
        if (j + i + 3u < (u32)CH) win[j + i + 3u] = __funnelshift_l(b3, b2, sh);
        a = b3;
        wi += 4u;
    }
    for (; i < n; i++) {
        const u32 b = payload_word(pw, n_words, ++wi);
        if (j + i < (u32)CH) win[j + i] = __funnelshift_l(b, a, sh);
        a = b;
    }
}

/* the rare inline macroblock whose code may exceed 64 bits: prefix, then the body word by word */
__device__ __noinline__ void put_long(BitPut &bp, u64 code, u32 meta, int num_refs, u32 *win, u32 wbase) {
    const u32 pre = long_prefix(code);
    bp_put(bp, (u64)pre, 63u - 2u * (u32)__clz((int)pre), win, wbase, CH);
    MbEntry e;
    int bl = encode_p16x16_body((int)(meta >> META_REF_SHIFT), long_dx(code), long_dy(code), num_refs, e);
    const u32 ws[3] = {e.w1, e.w2, e.w3};
    for (int i = 0; i < 3 && bl > 0; i++, bl -= 32) {
        const u32 n = (u32)min(bl, 32);
        bp_put(bp, (u64)(ws[i] >> (32u - n)), n, win, wbase, CH);
    }
}

/*
 * A spliced macroblock: mb_skip_run prefix, then its pre-encoded bits.  The partial words at both
 * ends go through the bit writer, the whole words in between are queued for the copy stage.
 */
__device__ __noinline__ void put_spliced(BitPut &bp, u64 code, ComposeSmem &sm, u32 wbase, u32 *n_copy) {
    const u32 *__restrict__ tab = sm.tab;
    const u32 *__restrict__ pw = sm.pay_words;
    const u32 n_words = sm.n_pay_words;
    u32 *win = sm.win;
    const u32 pre = dyn_prefix(code);
    bp_put(bp, (u64)pre, 63u - 2u * (u32)__clz((int)pre), win, wbase, CH);
    const u32 di = dyn_index(code);
    const u32 b0 = __ldg(tab + di), len = __ldg(tab + di + 1) - b0;
    const u32 hb = min(len, (32u - bp.fill) & 31u);
    if (hb) bp_put(bp, (u64)(payload_ld32(pw, n_words, b0) >> (32u - hb)), hb, win, wbase, CH);
    const u32 rest = len - hb, nw = rest >> 5, tb = rest & 31u;
    if (nw) {
        const u32 slot = atomicAdd(n_copy, 1u);
        if (slot < (u32)LCAP) sm.copy[slot] = make_uint4(bp.j - wbase, b0 + hb, nw, 0u);
        else splice_words(win, bp.j - wbase, nw, pw, n_words, b0 + hb);      /* list full: copy it here */
        bp.j += nw;
    }
    if (tb) bp_put(bp, (u64)(payload_ld32(pw, n_words, b0 + len - tb) >> (32u - tb)), tb, win, wbase, CH);
}

/* the copy stage: one lane per queued run */
__device__ __noinline__ void splice_queued(ComposeSmem &sm, u32 n_copy) {
    const u32 *__restrict__ pw = sm.pay_words;
    const u32 n_words = sm.n_pay_words;
    for (u32 d = threadIdx.x; d < n_copy; d += NT) {
        const uint4 c = sm.copy[d];
        splice_words(sm.win, c.x, c.z, pw, n_words, c.y);
    }
}

/* out of line: the backward scan is rare */
__device__ __noinline__ u32 zeros_before_win(const u32 *win, u32 b, u32 zrun_in) { return zeros_before(win, b, zrun_in); }

/* a group with insertions, the frame's last bytes, or one next to the slot's end: byte by byte */
__device__ __noinline__ void emit_bytes(uint8_t *fout, u32 cap, u32 pos, uint4 w, u32 ins, int vb) {
    const u32 ws[4] = {w.x, w.y, w.z, w.w};
    for (int k = 0; k < vb; k++) {
        if ((ins >> k) & 1u) { if (pos < cap) fout[pos] = 3; pos++; }
        if (pos < cap) fout[pos] = (uint8_t)(ws[k >> 2] >> (24 - 8 * (k & 3)));
        pos++;
    }
}

__global__ void __launch_bounds__(NT, COMPOSE_MIN_CTAS)
compose_frames_kernel(const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                      const H264GpuMbHint *__restrict__ hints,
                      const H264GpuDynDesc *__restrict__ dyns,
                      const u32 *__restrict__ dyn_tab,
                      const uint8_t *__restrict__ payload,
                      uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    __shared__ ComposeSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;

    {   /* the window is zero wherever nothing was written: cleared once here, then every flush clears what it used */
        uint4 *w4 = (uint4 *)sm.win;
        for (int i = tid; i < (CH + 8) / 4; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) { sm.n_copy[0] = 0; sm.n_copy[1] = 0; }
    }
    u32 spar = 0;                      /* which sm.scan buffer the next block scan uses */
    u32 cpar = 0;                      /* which sm.n_copy counter the current tile uses */
#if COMPOSE_TIMING
    long long stage_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, stage_t = clock64();
#endif

    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        /* ---- frame setup: the descriptor is read by every thread (one broadcast line) ------ */
        const uint4 *dq = (const uint4 *)(descs + f);
        const uint4 d2 = __ldg(dq + 2), d3 = __ldg(dq + 3);
        const u64 hint_off = (u64)d2.x | ((u64)d2.y << 32), out_off = (u64)d2.z | ((u64)d2.w << 32);
        const u32 cap = d3.x, dyn_id = d3.y;
        const int W = (int)(d3.z & 0xFFFFu), N = W * (int)(d3.z >> 16);
        const u32 hdr_total = 40u + (d3.w & 0xFFu);
        const int num_refs = (int)((d3.w >> 8) & 0xFFu);
        const int skip_spec = ((d3.w >> 16) & H264GPU_SKIP_SPEC) != 0;
        const int pred_spec = ((d3.w >> 16) & H264GPU_PRED_SPEC) != 0;
        const u32 nal_header = d3.w >> 24;
        const u32 w_magic = 0xFFFFFFFFu / (u32)W + 1u;         /* g / W == umulhi(g, magic): exact for g < 2^20, 1 < W < 2^10 */
        const uint2 *fh = (const uint2 *)hints + hint_off;
        const u32 *tab = nullptr;
        const u32 *pay_words = nullptr;
        u32 n_pay_words = 0;
        if (dyn_id != H264GPU_NO_DYN) {
            const uint4 *yq = (const uint4 *)(dyns + dyn_id);
            const uint4 y0 = __ldg(yq), y1 = __ldg(yq + 1);
            tab = dyn_tab + ((u64)y0.z | ((u64)y0.w << 32));
            pay_words = (const u32 *)(payload + ((u64)y0.x | ((u64)y0.y << 32)));
            n_pay_words = (y1.y + 3u) >> 2;
        }
        if (tid == 0) {                                        /* read again only after this frame's first barrier */
            sm.zrun[0] = 0; sm.zrun[1] = 0;
            sm.tab = tab; sm.pay_words = pay_words; sm.n_pay_words = n_pay_words;
            sm.cap = cap; sm.fout = out + out_off;
        }

        /* frame state every thread tracks identically (all of it follows from block-wide totals) */
        u32 bitpos = hdr_total;        /* slice bits composed so far (incl. the 40-bit Annex-B prefix) */
        int last_coded = -1;           /* frame index of the last non-skipped MB */
        u32 ins_total = 0;             /* emulation-prevention bytes inserted so far */
        u32 zpar = 0;                  /* which sm.zrun slot holds the current state */
        u32 wbase = 0;                 /* frame word held by window word 0; words below it have left for global memory */

        for (int t0 = 0; t0 < N; t0 += TILE) {
            const bool last = (t0 + TILE >= N);
            const bool tile0 = (t0 == 0);
            const int g0 = t0 + tid * MPT;
            const int nmb = min(max(N - g0, 0), MPT);          /* real macroblocks of this thread */

            /* ---- load: own hints, the row above, the left neighbour ---------- */
            uint2 own[MPT + 1];                                 /* own[0] = left neighbour of the first MB */
            uint2 abv[MPT + 2];                                 /* abv[j] = hint g0 - W - 1 + j */
#pragma unroll
            for (int k = 0; k <= MPT; k++) own[k] = make_uint2(0u, 0u);
#pragma unroll
            for (int j = 0; j < MPT + 2; j++) abv[j] = make_uint2(0u, 0u);
            if (nmb == MPT) {
                const uint4 *src = (const uint4 *)(fh + g0);    /* hint_off even, g0 % MPT == 0: 16-byte aligned */
#pragma unroll
                for (int k = 0; k < MPT / 2; k++) {
                    const uint4 v = __ldg(src + k);
                    own[1 + 2 * k] = make_uint2(v.x, v.y);
                    own[2 + 2 * k] = make_uint2(v.z, v.w);
                }
            } else {
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if (k < nmb) own[1 + k] = __ldg(fh + g0 + k);
            }
            if (nmb > 0) {
                if (g0 > 0) own[0] = __ldg(fh + g0 - 1);
                const int a0 = g0 - W - 1;
                if (a0 >= 0 && a0 + MPT + 2 <= N) {
#pragma unroll
                    for (int j = 0; j < MPT + 2; j++) abv[j] = __ldg(fh + a0 + j);
                } else {
#pragma unroll
                    for (int j = 0; j < MPT + 2; j++)
                        if (a0 + j >= 0 && a0 + j < N) abv[j] = __ldg(fh + a0 + j);
                }
            }

            /* ---- decide + code ------------------------------------------------ */
            u64 code[MPT];
            u32 meta[MPT];
            u32 dynlen = 0;                                     /* payload bits of this thread's spliced MBs */
            {
                int y = W == 1 ? g0 : (int)__umulhi((u32)g0, w_magic), x = g0 - y * W;
#pragma unroll
                for (int k = 0; k < MPT; k++) {
                    code[k] = 0; meta[k] = 0;
                    const uint2 h = own[1 + k];
                    const bool present = k < nmb;
                    const bool inter = ((h.y >> 8) & 0xFFu) == H264GPU_MB_INTER;
                    const int avA = x > 0, avB = y > 0;
                    const bool hasC = x + 1 < W;
                    const int avC = avB && (hasC || x > 0);
                    const uint2 c = hasC ? abv[k + 2] : abv[k];             /* C, else D (:391-399) */
                    /* uniform neighbourhood: A, B, C available and equal to this MB's (mv, ref, kind) */
                    const u32 hkey = h.y & 0xFFFFu;
                    const bool uni = inter && avA && avC &&
                                     own[k].x == h.x && abv[k + 1].x == h.x && c.x == h.x &&
                                     (own[k].y & 0xFFFFu) == hkey && (abv[k + 1].y & 0xFFFFu) == hkey && (c.y & 0xFFFFu) == hkey;
                    if (__all_sync(FULL, uni || !present)) {
                        if (present) meta[k] = decide_uniform(h.y, num_refs, skip_spec, code[k]);
                    } else if (present) {
                        if (inter) {
                            meta[k] = decide_inter(h.x, h.y, own[k].x, own[k].y, avA, abv[k + 1].x, abv[k + 1].y, avB,
                                                   c.x, c.y, avC, num_refs, pred_spec, skip_spec, code[k]);
                        } else {
                            const u32 di = h.y >> 16;
                            const u32 b0 = __ldg(tab + di), b1 = __ldg(tab + di + 1);
                            dynlen += b1 - b0;
                            prefetch_l1((const uint8_t *)pay_words + (b0 >> 3));      /* the payload is first touched here: */
                            prefetch_l1((const uint8_t *)pay_words + (b1 >> 3));      /* have it on its way before the splice */
                            meta[k] = META_DYN | META_CODED;
                            code[k] = (u64)di << 48;
                        }
                    }
                    if (++x == W) { x = 0; y++; }
                }
            }

            STAGE_MARK(0);                                      /* load + decide */
            if (g0 + TILE < N && (tid & 3) == 0) prefetch_l2(fh + g0 + TILE);   /* next tile's hints: one 128-byte line per 4 threads */

            /* ---- mb_skip_run prefixes ------------------------------------------ */
            int last_coded_total = last_coded;
            if (skip_spec) {
                int lc = -1;
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if (meta[k] & META_CODED) lc = g0 + k;
                int blk_max;
                int prev = max(block_excl_scan_max(lc, (int *)sm.scan[spar], blk_max), last_coded);
                spar ^= 1u;
                last_coded_total = max(blk_max, last_coded);
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if (meta[k] & META_CODED) {
                        const u32 pre = (u32)(g0 + k - prev);              /* run + 1 */
                        attach_prefix(code[k], meta[k], pre, 63 - 2 * __clz((int)pre));
                        prev = g0 + k;
                    }
            } else {
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if (meta[k] & META_CODED) attach_prefix(code[k], meta[k], 1u, 1);   /* ue(0), :630 */
            }

            /* ---- merge: consecutive plain codes -> one <= 64-bit item per run, bit count of the thread ---- */
            u32 L = dynlen;
            {
                u64 acc = 0;
                u32 accn = 0;
#pragma unroll
                for (int k = 0; k < MPT; k++) {
                    const u64 c = code[k];
                    const u32 m = meta[k];
                    L += m & META_LEN_MASK;
                    code[k] = 0; meta[k] = 0;
                    if (m & (META_DYN | META_LONG)) {
                        if (k > 0 && accn) { code[k - 1] = acc; meta[k - 1] = accn; acc = 0; accn = 0; }
                        code[k] = c; meta[k] = m;
                    } else if (m & META_CODED) {
                        u64 sc = 0; u32 sn = 0;
                        const bool seal = merge_plain(acc, accn, c, m & META_LEN_MASK, sc, sn);
                        if (k > 0 && seal) { code[k - 1] = sc; meta[k - 1] = sn; }
                    }
                }
                if (accn) { code[MPT - 1] = acc; meta[MPT - 1] = accn; }
            }

            /* ---- slice trailer: trailing mb_skip_run + rbsp_stop_one_bit ---------- */
            const bool owns_trailer = last && nmb > 0 && g0 + nmb == N;
            u32 tr_code = 1u, tr_len = 1u;
            if (owns_trailer && skip_spec) {
                const u32 run_end = (u32)(N - 1 - last_coded_total);
                if (run_end > 0u) {
                    const int pn = 63 - 2 * __clz((int)(run_end + 1u));
                    tr_code = ((run_end + 1u) << 1) | 1u;
                    tr_len = (u32)pn + 1u;
                }
            }
            if (owns_trailer) L += tr_len;

            /* ---- bit offsets: prefix sum of the threads' bit counts ----------------- */
            u32 total;
            int tile_has_dyn;
            const u32 O = bitpos + block_excl_scan(L, sm.scan[spar], total, dynlen != 0u, tile_has_dyn);
            spar ^= 1u;
            const u32 E = O + L;
            const u32 end = bitpos + total;
            const u32 total_bytes = (end + 7u) >> 3;                       /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* frame words [.., w_stop) are complete after this tile */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end; /* last word this tile writes to */
            const u32 warp_end = __shfl_sync(FULL, E, 31);
            STAGE_MARK(1);                                      /* prefixes + merge + scan (incl. its barrier) */
#if COMPOSE_ABLATE == 1                                                     /* timing experiment: stop after decide + scan */
            {
                u32 x = O;
#pragma unroll
                for (int k = 0; k < MPT; k++) x ^= (u32)code[k] ^ (u32)(code[k] >> 32) ^ meta[k];
                if (x == 0x12345u) sizes[f] = x;
                bitpos = end; last_coded = last_coded_total;
                continue;
            }
#endif

            for (;;) {
                /* ---- write: sequential bit writer per thread, clipped to the window ---------- */
                BitPut bp;
                bp_init(bp, (tile0 && tid == 0) ? 0u : O);
                const u32 key = bp.j;
                if (tile0 && tid == 0) {
                    const uint4 h0 = __ldg(dq), h1 = __ldg(dq + 1);
                    const u64 hdr[4] = {(u64)h0.x | ((u64)h0.y << 32), (u64)h0.z | ((u64)h0.w << 32),
                                        (u64)h1.x | ((u64)h1.y << 32), (u64)h1.z | ((u64)h1.w << 32)};
                    for (u32 p = 0; p < hdr_total; p += 32u) {
                        const u32 n = min(32u, hdr_total - p);
                        bp_put(bp, (u64)(header_bits32(hdr, nal_header, p) >> (32u - n)), n, sm.win, wbase, CH);
                    }
                }
#pragma unroll
                for (int k = 0; k < MPT; k++) {
                    const u32 m = meta[k];
                    if (m) {
                        if (!(m & (META_DYN | META_LONG))) bp_put(bp, code[k], m, sm.win, wbase, CH);
                        else if (m & META_DYN) put_spliced(bp, code[k], sm, wbase, &sm.n_copy[cpar]);
                        else put_long(bp, code[k], m, num_refs, sm.win, wbase);
                    }
                }
                if (owns_trailer) bp_put(bp, (u64)tr_code, tr_len, sm.win, wbase, CH);
                u32 T = 0, tj = 0;
                bool has_tail = false;
                if (bp.fill) {
                    if (bp.head) bp.H = bp.acc;
                    else { T = bp.acc; tj = bp.j; has_tail = true; }
                }

                /* ---- merge the partial words of neighbouring threads (segmented suffix OR) ---- */
                {
                    u32 X = bp.H;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const u32 v = __shfl_down_sync(FULL, X, d);
                        const u32 kk = __shfl_down_sync(FULL, key, d);
                        if (lane + d < 32 && kk == key) X |= v;
                    }
                    const u32 Xn = __shfl_down_sync(FULL, X, 1);
                    const u32 kn = __shfl_down_sync(FULL, key, 1);
                    if (has_tail && tj - wbase < (u32)CH) {
                        u32 v = T;
                        if (lane < 31 && kn == tj) v |= Xn;
                        if ((warp_end >> 5) == tj) atomicOr(&sm.win[tj - wbase], v);   /* later warps (or the next tile) add to this word */
                        else sm.win[tj - wbase] = v;
                    }
                    if (lane == 0 && X != 0u && key - wbase < (u32)CH) atomicOr(&sm.win[key - wbase], X);
                }

#if COMPOSE_ABLATE == 2                                                     /* timing experiment: stop after the bit writer */
                __syncthreads();
                if (sm.win[(tid * 7) % CH] == 0x12345u) sizes[f] = 1;
                {
                    uint4 *w4 = (uint4 *)sm.win;
                    for (u32 i = (u32)tid; i < (CH + 3u) / 4u; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
                }
                if (tid == 0) { sm.n_copy[0] = 0; sm.n_copy[1] = 0; }
                __syncthreads();
                break;
#endif
                STAGE_MARK(2);                                  /* bit writer + partial-word merge */
                const bool fits = w_last - wbase < (u32)CH;                /* the tile's last word lies inside the window */
                const bool flush = last || !fits;
                bool synced = false;

                /* ---- splice: whole words of spliced MBs, one lane per queued run ----------- */
                if (tile_has_dyn) {
                    __syncthreads();
                    synced = true;
                    splice_queued(sm, min(sm.n_copy[cpar], (u32)LCAP));
                    cpar ^= 1u;
                    if (tid == 0) sm.n_copy[cpar] = 0;                     /* nobody reads that one before the next tile's barrier */
                }
                STAGE_MARK(3);                                  /* splice copy (incl. its barrier) */
                if (!flush) break;                                         /* keep filling the window with the next tile */

                /* ---- emit: emulation prevention + store of frame words [wbase, e_hi) --------- */
                if (!synced || tile_has_dyn) __syncthreads();
                STAGE_MARK(4);                                  /* barrier before emit */
                const u32 e_hi = fits ? w_stop : wbase + CH;
                const u32 chunk_byte0 = wbase << 2;
                const int chunk_bytes = e_hi > wbase ? (int)((fits ? min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0;
                const u32 zrun_in = sm.zrun[zpar];
                const u32 cap = sm.cap;
                uint8_t *fout = sm.fout;
                u32 *fout32 = (u32 *)fout;
                for (int q0 = 0; q0 * 16 < chunk_bytes; q0 += NT) {
                    const int q = q0 + tid;
                    const int vb = min(max(chunk_bytes - 16 * q, 0), 16);
                    const int vb_next = tid == NT - 1 ? 0 : min(max(chunk_bytes - 16 * (q + 1), 0), 16);
                    uint4 w = make_uint4(FULL, FULL, FULL, FULL);
                    u32 wn = 0, ins = 0;
                    if (vb > 0) {
                        w = ((const uint4 *)sm.win)[q]; wn = sm.win[4 * q + 4];
                        const u32 hz = ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) |
                                       ((w.z - 0x01010101u) & ~w.z) | ((w.w - 0x01010101u) & ~w.w);
                        if ((hz & 0x80808080u) || (w.x >> 26) == 0u) {
                            /* a zero byte, or a first byte <= 3 that may close a zero run of the bytes before */
                            const u32 valid = (1u << vb) - 1u;
                            u32 z = (zero_nib(w.x) | (zero_nib(w.y) << 4) | (zero_nib(w.z) << 8) | (zero_nib(w.w) << 12)) & valid;
                            u32 keep = valid;
                            if (chunk_byte0 == 0u && q == 0) { z &= ~0x1Fu; keep &= ~0x1Fu; }   /* start code + NAL header */
                            const bool need = ((w.x >> 26) == 0u) && (keep & 1u);
                            if (need || (z & (z << 1))) {
                                u32 rin = 0;
                                if (need && (q == 0 || (sm.win[4 * q - 1] & 0xFFu) == 0u)) rin = zeros_before_win(sm.win, 16u * (u32)q, zrun_in);
                                const u32 m3 = zero_nib(w.x & 0xFCFCFCFCu) | (zero_nib(w.y & 0xFCFCFCFCu) << 4) |
                                               (zero_nib(w.z & 0xFCFCFCFCu) << 8) | (zero_nib(w.w & 0xFCFCFCFCu) << 12);
                                ins = epb_group_mask(z, m3 & keep, vb, rin);
                            }
                        }
                        /* the holder of the window's last byte hands the zero run to the next window */
                        if (16 * (q + 1) >= chunk_bytes)
                            sm.zrun[zpar ^ 1u] = zeros_before_win(sm.win, (u32)chunk_bytes, zrun_in);
                    }
                    const u32 cnt = (u32)__popc(ins);
                    sm.gi[tid] = (cnt << 8) | (u32)vb;
                    u32 before = ins_total, pass_ins = 0;
                    int any_ins;
                    before += block_excl_scan(cnt, sm.scan[spar], pass_ins, 0, any_ins);
                    spar ^= 1u;
                    /* the whole pass stays clear of the slot's end, else everybody goes byte by byte */
                    const bool roomy = chunk_byte0 + 16u * (u32)(q0 + NT) + ins_total + pass_ins + 4u <= cap;
                    if (vb > 0) {
                        const u32 D = chunk_byte0 + 16u * (u32)q + before;      /* output byte of the group's first byte */
                        const bool clean = cnt == 0u && (vb & 3) == 0 && roomy;
                        bool prev_clean = tid > 0 && roomy, next_clean = tid < NT - 1 && vb_next == 16 && roomy;
                        if (pass_ins) {
                            if (tid > 0) prev_clean = prev_clean && sm.gi[tid - 1] == 16u;
                            if (tid < NT - 1) next_clean = next_clean && sm.gi[tid + 1] == 16u;
                        }
                        if (clean) {
                            /* vb / 4 whole source words leave as re-aligned 32-bit words */
                            const u32 s = (4u - (D & 3u)) & 3u;                 /* source bytes before the first aligned word */
                            const u32 sel = realign_selector(s);
                            const int nw = vb >> 2;
                            u32 *dst = fout32 + ((D + s) >> 2);
                            const u32 o0 = __byte_perm(w.x, w.y, sel), o1 = __byte_perm(w.y, w.z, sel),
                                      o2 = __byte_perm(w.z, w.w, sel), o3 = __byte_perm(w.w, wn, sel);
                            const bool whole_last = s == 0u || (vb == 16 && next_clean);   /* else the last word keeps to its own bytes */
                            if (nw > 1 || whole_last) dst[0] = o0;
                            if (nw > 2 || (nw == 2 && whole_last)) dst[1] = o1;
                            if (nw > 3 || (nw == 3 && whole_last)) dst[2] = o2;
                            if (nw == 4 && whole_last) dst[3] = o3;
                            if (!whole_last) {
                                const u32 ol = nw == 4 ? o3 : nw == 3 ? o2 : nw == 2 ? o1 : o0;
                                const u32 base = D + s + 4u * (u32)(nw - 1);
                                for (u32 k = 0; k < 4u - s; k++) fout[base + k] = (uint8_t)(ol >> (8u * k));
                            }
                            if (s && (!prev_clean || vb < 16))                  /* nobody else covers the bytes before the first aligned word */
                                for (u32 k = 0; k < s; k++) fout[D + k] = (uint8_t)(w.x >> (24u - 8u * k));
                        } else {
                            emit_bytes(fout, cap, D, w, ins, vb);
                        }
                    }
                    ins_total += pass_ins;
                }
                if (chunk_bytes > 0) zpar ^= 1u;
                STAGE_MARK(5);                                  /* emit passes */
                __syncthreads();                                               /* window fully consumed */
                {   /* clear the words this window used */
                    const u32 used = min((u32)CH, w_last - wbase + 1u);
                    uint4 *w4 = (uint4 *)sm.win;
                    for (u32 i = (u32)tid; i < (used + 3u) / 4u; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
                }
                STAGE_MARK(6);                                  /* barrier + window clear */
                if (fits) break;                                               /* (frame done) */
                wbase += CH;                                                   /* a full window has left; the tile goes on in the next one */
                __syncthreads();
            }

            bitpos = end;
            last_coded = last_coded_total;
            if (last && tid == 0) sizes[f] = total_bytes + ins_total;
        }
    }
#if COMPOSE_TIMING
    if (tid == 0)
        for (int i = 0; i < 8; i++) atomicAdd(&g_stage_cycles[i], (unsigned long long)stage_acc[i]);
#endif
}

#if COMPOSE_TIMING
cudaError_t read_stage_cycles(unsigned long long *h8) {
    cudaError_t e = cudaMemcpyFromSymbol(h8, g_stage_cycles, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_stage_cycles, z, sizeof z);
    return e;
}
#endif

/* ------------------------------------------------------------ rasteriser */

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_MAX_REGIONS = H264GPU_MAX_REGIONS;

__global__ void __launch_bounds__(RASTER_THREADS)
raster_regions_kernel(const H264GpuRasterDesc *__restrict__ rdescs, int n_frames,
                      const H264GpuRegion *__restrict__ regions,
                      const H264GpuMbHint *__restrict__ dyn_info,
                      H264GpuMbHint *__restrict__ hints) {
    __shared__ uint4 s_reg[RASTER_MAX_REGIONS];
    const int f = blockIdx.x;
    if (f >= n_frames) return;
    const H264GpuRasterDesc rd = rdescs[f];
    const int W = rd.mb_w, N = W * (int)rd.mb_h;
    const int nreg = min((int)rd.region_count, RASTER_MAX_REGIONS);
    for (int r = threadIdx.x; r < nreg; r += RASTER_THREADS)
        s_reg[r] = __ldg((const uint4 *)(regions + rd.region_first) + r);
    __syncthreads();
    uint2 *dst = (uint2 *)hints + rd.hint_off;
    const uint2 *dinfo = (const uint2 *)dyn_info + rd.dyn_info_off;
    for (int g = blockIdx.y * RASTER_THREADS + threadIdx.x; g < N; g += gridDim.y * RASTER_THREADS) {
        const int y = g / W, x = g - y * W;
        uint2 h = make_uint2(0u, 0u);                      /* uncovered: (ref 0, mv 0) */
        for (int r = nreg - 1; r >= 0; r--) {              /* later regions occlude earlier ones */
            const uint4 q = s_reg[r];
            const int x0 = (int16_t)(q.x & 0xFFFF), y0 = (int16_t)(q.x >> 16);
            const int x1 = (int16_t)(q.y & 0xFFFF), y1 = (int16_t)(q.y >> 16);
            if (x >= x0 && x < x1 && y >= y0 && y < y1) {
                const u32 kind = (q.w >> 8) & 0xFFu;
                if (kind == H264GPU_MB_INTER) h = make_uint2(q.z, q.w & 0xFFFFu);
                else {
                    const u32 local = (u32)((y - y0) * (x1 - x0) + (x - x0));
                    h = __ldg(dinfo + (q.w >> 16) + local);
                }
                break;
            }
        }
        dst[g] = h;
    }
}

/* ---------------------------------------------------------- stream packing */

constexpr int PACK_SCAN_THREADS = 1024;

/* offsets[f] = sum of min(size, cap) over frames < f; single CTA, chunked */
__global__ void __launch_bounds__(PACK_SCAN_THREADS)
pack_offsets_kernel(const H264GpuFrameDesc *__restrict__ descs, const u32 *__restrict__ sizes,
                    int n_frames, unsigned long long *__restrict__ offsets) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_frames; c0 += PACK_SCAN_THREADS) {
        const int f = c0 + tid;
        unsigned long long v = 0;
        if (f < n_frames) v = min(sizes[f], descs[f].out_cap);
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long base = s_base;
        const unsigned long long excl = base + s_warp[warp] + inc - v;
        if (f < n_frames) offsets[f] = excl;
        __syncthreads();
        if (tid == PACK_SCAN_THREADS - 1) s_base = excl + v;
        __syncthreads();
    }
    if (tid == 0) offsets[n_frames] = s_base;
}

/* copy slot f to packed[offsets[f] ..) with aligned 32-bit stores (byte-permute realignment) */
__global__ void __launch_bounds__(256)
pack_copy_kernel(const H264GpuFrameDesc *__restrict__ descs, const u32 *__restrict__ sizes, int n_frames,
                 const uint8_t *__restrict__ slots, uint8_t *__restrict__ packed,
                 const unsigned long long *__restrict__ offsets) {
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const u32 n = min(sizes[f], descs[f].out_cap);
        const u32 *src = (const u32 *)(slots + descs[f].out_off);     /* 16-byte aligned */
        const unsigned long long d0 = offsets[f];
        uint8_t *dst = packed + d0;
        const u32 mis = (u32)(d0 & 3ull);                               /* dst misalignment */
        const u32 head = mis ? min(4u - mis, n) : 0u;                   /* bytes before the first aligned dst word */
        if (threadIdx.x < head) dst[threadIdx.x] = slots[descs[f].out_off + threadIdx.x];
        const u32 body = (n - head) >> 2;                               /* aligned dst words */
        u32 *dst32 = (u32 *)(dst + head);
        const u32 n_src_words = (n + 3u) >> 2;
        for (u32 k = threadIdx.x; k < body; k += blockDim.x) {
            /* dst word k holds src bytes [head + 4k, head + 4k + 4) */
            const u32 sw = (head + 4u * k) >> 2, sh = head & 3u;
            const u32 a = __ldg(src + sw);
            const u32 b = (sh && sw + 1u < n_src_words) ? __ldg(src + sw + 1u) : 0u;
            dst32[k] = __byte_perm(a, b, 0x3210u + 0x1111u * sh);
        }
        const u32 tail0 = head + 4u * body;
        if (threadIdx.x < n - tail0) dst[tail0 + threadIdx.x] = slots[descs[f].out_off + tail0 + threadIdx.x];
    }
}

/* ---------------------------------------------------------------- launch */

cudaError_t launch_compose(const H264GpuFrameDesc *d_desc, int n_frames,
                           const H264GpuMbHint *d_hints, const H264GpuDynDesc *d_dyn,
                           const uint32_t *d_dyn_tab, const uint8_t *d_payload,
                           uint8_t *d_out, uint32_t *d_sizes, int sm_count, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    static int ctas_per_sm = 0;           /* resident CTAs per SM for this build (registers / shared memory) */
    if (ctas_per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, compose_frames_kernel, NT, 0) != cudaSuccess || n < 1) n = 1;
        ctas_per_sm = n;
    }
    const int resident = sm_count * ctas_per_sm;      /* persistent grid: a multiple of the SM count */
    int grid = n_frames < resident ? n_frames : resident;
    compose_frames_kernel<<<grid, NT, 0, stream>>>(d_desc, n_frames, d_hints, d_dyn, d_dyn_tab,
                                                   d_payload, d_out, d_sizes);
    return cudaGetLastError();
}

cudaError_t launch_raster(const H264GpuRasterDesc *d_rdesc, int n_frames, int max_mbs,
                          const H264GpuRegion *d_regions, const H264GpuMbHint *d_dyn_info,
                          H264GpuMbHint *d_hints, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    int bx = (max_mbs + RASTER_THREADS * 4 - 1) / (RASTER_THREADS * 4);
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dim3 grid(n_frames, bx);
    raster_regions_kernel<<<grid, RASTER_THREADS, 0, stream>>>(d_rdesc, n_frames, d_regions, d_dyn_info, d_hints);
    return cudaGetLastError();
}

cudaError_t launch_pack(const H264GpuFrameDesc *d_desc, const uint32_t *d_sizes, int n_frames,
                        const uint8_t *d_slots, uint8_t *d_packed, unsigned long long *d_offsets,
                        cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    pack_offsets_kernel<<<1, PACK_SCAN_THREADS, 0, stream>>>(d_desc, d_sizes, n_frames, d_offsets);
    int grid = n_frames < 148 * 8 ? n_frames : 148 * 8;
    pack_copy_kernel<<<grid, 256, 0, stream>>>(d_desc, d_sizes, n_frames, d_slots, d_packed, d_offsets);
    return cudaGetLastError();
}

} /* namespace h264scroll */
