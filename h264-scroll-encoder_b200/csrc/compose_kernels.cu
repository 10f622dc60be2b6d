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
constexpr int MPT = 8;                   /* consecutive macroblocks per thread */
constexpr int TILE = NT * MPT;           /* macroblocks per tile */
constexpr int NW = NT / 32;
constexpr int CH = 6144;                 /* window: slice words staged per pass (24 KB) */
constexpr int LCAP = 512;                /* splice copy list entries per pass */
constexpr u32 FULL = 0xFFFFFFFFu;

struct FrameState {
    u32 bitpos;        /* slice bits emitted so far (incl. the 40-bit Annex-B prefix) */
    u32 carry;         /* partial slice word at bitpos >> 5 */
    int last_coded;    /* frame index of the last non-skipped MB, -1 if none */
    u32 zrun;          /* zero bytes closing the bytes emitted so far (rbsp_to_ebsp state) */
    u32 ins_total;     /* emulation-prevention bytes inserted so far */
    u32 n_copy;        /* entries in the splice copy list */
    u32 zrun_next;
};

struct __align__(16) ComposeSmem {
    u32 win[CH + 8];                     /* slice bits of the pass, MSB-first words (+ read-ahead pad) */
    uint4 copy[LCAP];                    /* {first frame word, payload bit, words, -} */
    u32 tr[NT];                          /* emit: zero bytes closing each 16-byte group */
    u32 gi[NT];                          /* emit: (insert count << 8) | valid bytes of each group */
    u32 scan[NW + 1];
    H264GpuFrameDesc desc;
    H264GpuDynDesc dyn;
    FrameState st;
};

/* exclusive prefix sum over the CTA; every thread gets its offset and the total */
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *s_scan, u32 &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    const u32 wsum = lane < NW ? s_scan[lane] : 0u;    /* every warp scans the NW warp totals itself */
    u32 winc = wsum;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, winc, d);
        if (lane >= d) winc += t;
    }
    const u32 wexc = __shfl_sync(FULL, winc - wsum, warp);
    total = __shfl_sync(FULL, winc, NW - 1);
    __syncthreads();                                   /* s_scan may be reused by the caller */
    return wexc + inc - v;
}

/* exclusive prefix max over the CTA (identity -1 as int); `total` = max over all threads */
__device__ __forceinline__ int block_excl_scan_max(int v, int *s_scan, int &total) {
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
    __syncthreads();
    int exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = -1;
    return max(wexc, exc);
}

/* 32 payload bits from bit position p (payload words are big-endian in memory) */
__device__ __forceinline__ u32 payload_ld32(const u32 *__restrict__ pw, u32 n_words, u32 p) {
    const u32 wi = p >> 5;
    const u32 a = wi < n_words ? bswap32(__ldg(pw + wi)) : 0u;
    const u32 b = wi + 1u < n_words ? bswap32(__ldg(pw + wi + 1u)) : 0u;
    return __funnelshift_l(b, a, p & 31u);
}

__global__ void __launch_bounds__(NT, 3)
compose_frames_kernel(const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                      const H264GpuMbHint *__restrict__ hints,
                      const H264GpuDynDesc *__restrict__ dyns,
                      const u32 *__restrict__ dyn_tab,
                      const uint8_t *__restrict__ payload,
                      uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    __shared__ ComposeSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;

    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        /* ---- frame setup ------------------------------------------------ */
        __syncthreads();
        if (tid < (int)(sizeof(H264GpuFrameDesc) / 4))
            ((u32 *)&sm.desc)[tid] = ((const u32 *)&descs[f])[tid];
        __syncthreads();
        const int W = sm.desc.mb_w, N = W * (int)sm.desc.mb_h;
        const u32 w_magic = 0xFFFFFFFFu / (u32)W + 1u;         /* g / W == umulhi(g, magic): exact for g < 2^20, 1 < W < 2^10 */
        const int num_refs = sm.desc.num_refs;
        const int skip_spec = (sm.desc.flags & H264GPU_SKIP_SPEC) != 0;
        const int pred_spec = (sm.desc.flags & H264GPU_PRED_SPEC) != 0;
        const u32 cap = sm.desc.out_cap;
        const u32 nal_header = sm.desc.nal_header;
        const u32 hdr_total = 40u + sm.desc.hdr_nbits;
        const u32 dyn_id = sm.desc.dyn_id;
        if (dyn_id != H264GPU_NO_DYN && tid < (int)(sizeof(H264GpuDynDesc) / 4))
            ((u32 *)&sm.dyn)[tid] = ((const u32 *)&dyns[dyn_id])[tid];
        if (tid == 0) {
            sm.st.bitpos = hdr_total;
            sm.st.carry = 0;
            sm.st.last_coded = -1;
            sm.st.zrun = 0;
            sm.st.ins_total = 0;
            sm.st.n_copy = 0;
        }
        __syncthreads();
        const uint2 *fh = (const uint2 *)hints + sm.desc.hint_off;
        const u32 *tab = nullptr;
        const u32 *pay_words = nullptr;
        u32 n_pay_words = 0;
        if (dyn_id != H264GPU_NO_DYN) {
            tab = dyn_tab + sm.dyn.table_off;
            pay_words = (const u32 *)(payload + sm.dyn.payload_off);
            n_pay_words = (sm.dyn.payload_bytes + 3u) >> 2;
        }
        uint8_t *fout = out + sm.desc.out_off;
        u32 *fout32 = (u32 *)fout;

        for (int t0 = 0; t0 < N; t0 += TILE) {
            const bool last = (t0 + TILE >= N);
            const bool tile0 = (t0 == 0);
            const int g0 = t0 + tid * MPT;
            const int nmb = min(max(N - g0, 0), MPT);          /* real macroblocks of this thread */

            /* the window is cleared while the hint loads are in flight */
            {
                uint4 *w4 = (uint4 *)sm.win;
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                for (int i = tid; i < (CH + 8) / 4; i += NT) w4[i] = z4;
                if (tid == 0 && !tile0) sm.win[0] = sm.st.carry;   /* bits of the previous tile in the shared word */
            }

            /* ---- load: own hints, the row above, the left neighbour ---------- */
            uint2 own[MPT + 1];                                 /* own[0] = left neighbour of the first MB */
            uint2 abv[MPT + 2];                                 /* abv[j] = hint g0 - W - 1 + j */
#pragma unroll
            for (int k = 0; k <= MPT; k++) own[k] = make_uint2(0u, 0u);
#pragma unroll
            for (int j = 0; j < MPT + 2; j++) abv[j] = make_uint2(0u, 0u);
            if (nmb == MPT) {
                const uint4 *src = (const uint4 *)(fh + g0);    /* hint_off even, g0 % 8 == 0: 16-byte aligned */
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
#pragma unroll
                for (int j = 0; j < MPT + 2; j++)
                    if (a0 + j >= 0 && a0 + j < N) abv[j] = __ldg(fh + a0 + j);
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
                    if (k < nmb) {
                        const uint2 h = own[1 + k];
                        const u32 hk = (h.y >> 8) & 0xFFu;
                        if (hk == H264GPU_MB_INTER) {
                            const int avA = x > 0, avB = y > 0;
                            const bool hasC = x + 1 < W;
                            const int avC = avB && (hasC || x > 0);
                            const uint2 c = hasC ? abv[k + 2] : abv[k];     /* C, else D (:391-399) */
                            meta[k] = decide_inter(h.x, h.y, own[k].x, own[k].y, avA, abv[k + 1].x, abv[k + 1].y, avB,
                                                   c.x, c.y, avC, num_refs, pred_spec, skip_spec, code[k]);
                        } else {
                            const u32 di = h.y >> 16;
                            const u32 b0 = __ldg(tab + di), b1 = __ldg(tab + di + 1);
                            dynlen += b1 - b0;
                            meta[k] = META_DYN | META_CODED;
                            code[k] = (u64)di << 48;
                        }
                    }
                    if (++x == W) { x = 0; y++; }
                }
            }

            /* ---- mb_skip_run prefixes ------------------------------------------ */
            int prev_in = -1, last_coded_total = -1;
            if (skip_spec) {
                int lc = -1;
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if (meta[k] & META_CODED) lc = g0 + k;
                prev_in = block_excl_scan_max(lc, (int *)sm.scan, last_coded_total);
                prev_in = max(prev_in, sm.st.last_coded);
                last_coded_total = max(last_coded_total, sm.st.last_coded);
                int prev = prev_in;
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

            /* ---- bit offsets: prefix sum of the threads' bit counts ----------------- */
            u32 L = dynlen + (owns_trailer ? tr_len : 0u);
#pragma unroll
            for (int k = 0; k < MPT; k++) L += meta[k] & META_LEN_MASK;
            const u32 bitpos = sm.st.bitpos;
            u32 total;
            const u32 O = bitpos + block_excl_scan(L, sm.scan, total);     /* the scan's barriers also publish the cleared window */
            const u32 E = O + L;
            const u32 end = bitpos + total;
            const u32 w_lo = bitpos >> 5;
            const u32 total_bytes = (end + 7u) >> 3;                       /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* words [.., w_stop) are emitted */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end;
            const u32 warp_end = __shfl_sync(FULL, E, 31);
            bool first_pass = true;

            for (u32 wbase = tile0 ? 0u : w_lo; wbase <= w_last; wbase += CH) {
                if (!first_pass) {
                    __syncthreads();
                    uint4 *w4 = (uint4 *)sm.win;
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                    for (int i = tid; i < (CH + 8) / 4; i += NT) w4[i] = z4;
                    if (tid == 0) sm.st.n_copy = 0;
                    __syncthreads();
                }
                first_pass = false;

                /* ---- write: sequential bit writer per thread ------------------------- */
                BitPut bp;
                bp_init(bp, (tile0 && tid == 0) ? 0u : O);
                const u32 key = bp.j;
                if (tile0 && tid == 0) {
                    for (u32 p = 0; p < hdr_total; p += 32u) {
                        const u32 n = min(32u, hdr_total - p);
                        bp_put(bp, (u64)(header_bits32(sm.desc.hdr_bits, nal_header, p) >> (32u - n)), n, sm.win, wbase, CH);
                    }
                }
                {
                    int prev = prev_in;
#pragma unroll
                    for (int k = 0; k < MPT; k++) {
                        const u32 m = meta[k];
                        if (m & META_CODED) {
                            const u32 n = m & META_LEN_MASK;
                            if (!(m & (META_DYN | META_LONG | META_SPLIT))) {
                                bp_put(bp, code[k], n, sm.win, wbase, CH);
                            } else {
                                const u32 pre = skip_spec ? (u32)(g0 + k - prev) : 1u;
                                const u32 pn = 63u - 2u * (u32)__clz((int)pre);
                                bp_put(bp, (u64)pre, pn, sm.win, wbase, CH);
                                if (m & META_SPLIT) {
                                    bp_put(bp, code[k], n - pn, sm.win, wbase, CH);
                                } else if (m & META_LONG) {
                                    MbEntry e;
                                    int bl = encode_p16x16_body((int)(m >> META_REF_SHIFT), long_dx(code[k]), long_dy(code[k]), num_refs, e);
                                    bp_put(bp, (u64)e.w1, 32u, sm.win, wbase, CH);
                                    bp_put(bp, (u64)e.w2, 32u, sm.win, wbase, CH);
                                    bl -= 64;
                                    bp_put(bp, (u64)(e.w3 >> (32 - bl)), (u32)bl, sm.win, wbase, CH);
                                } else {
                                    /* spliced MB: partial words through the writer, whole words through the copy list */
                                    const u32 di = (u32)(code[k] >> 48);
                                    const u32 b0 = __ldg(tab + di), len = __ldg(tab + di + 1) - b0;
                                    const u32 hb = min(len, (32u - bp.fill) & 31u);
                                    if (hb) bp_put(bp, (u64)(payload_ld32(pay_words, n_pay_words, b0) >> (32u - hb)), hb, sm.win, wbase, CH);
                                    const u32 rest = len - hb, nw = rest >> 5, tb = rest & 31u;
                                    if (nw) {
                                        const u32 slot = atomicAdd(&sm.st.n_copy, 1u);
                                        if (slot < (u32)LCAP) sm.copy[slot] = make_uint4(bp.j, b0 + hb, nw, 0u);
                                        else
                                            for (u32 i = 0; i < nw; i++)
                                                if (bp.j + i - wbase < (u32)CH)
                                                    sm.win[bp.j + i - wbase] = payload_ld32(pay_words, n_pay_words, b0 + hb + 32u * i);
                                        bp.j += nw;
                                    }
                                    if (tb) bp_put(bp, (u64)(payload_ld32(pay_words, n_pay_words, b0 + len - tb) >> (32u - tb)), tb, sm.win, wbase, CH);
                                }
                            }
                            prev = g0 + k;
                        }
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
                        if ((warp_end >> 5) == tj) atomicOr(&sm.win[tj - wbase], v);   /* later warps add to this word */
                        else sm.win[tj - wbase] = v;
                    }
                    if (lane == 0 && X != 0u && key - wbase < (u32)CH) atomicOr(&sm.win[key - wbase], X);
                }
                __syncthreads();

                /* ---- splice: whole words of spliced MBs, 8 lanes per entry ------------------ */
                {
                    const u32 n_copy = min(sm.st.n_copy, (u32)LCAP);
                    const u32 sub = (u32)tid & 7u;
                    for (u32 d = (u32)tid >> 3; d < n_copy; d += NT / 8) {
                        const uint4 c = sm.copy[d];
                        for (u32 i = sub; i < c.z; i += 8u) {
                            const u32 j = c.x + i - wbase;
                            if (j < (u32)CH) sm.win[j] = payload_ld32(pay_words, n_pay_words, c.y + 32u * i);
                        }
                    }
                    if (n_copy) __syncthreads();
                }

                /* ---- emit: emulation prevention + store of words [wbase, e_hi) -------------- */
                const u32 e_hi = min(wbase + CH, w_stop);
                const u32 chunk_byte0 = wbase << 2;
                const int chunk_bytes = e_hi > wbase ? (int)((last ? min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0;
                for (int q0 = 0; q0 * 16 < chunk_bytes; q0 += NT) {
                    const int q = q0 + tid;
                    const int vb = min(max(chunk_bytes - 16 * q, 0), 16);
                    const int vb_next = tid == NT - 1 ? 0 : min(max(chunk_bytes - 16 * (q + 1), 0), 16);
                    uint4 w = make_uint4(FULL, FULL, FULL, FULL);
                    u32 wn = 0;
                    if (vb > 0) { w = ((const uint4 *)sm.win)[q]; wn = sm.win[4 * q + 4]; }
                    u32 z = zero_nib(w.x) | (zero_nib(w.y) << 4) | (zero_nib(w.z) << 8) | (zero_nib(w.w) << 12);
                    const u32 valid = (1u << vb) - 1u;
                    z &= valid;
                    u32 keep = valid;                                      /* bytes that may take an insertion */
                    if (chunk_byte0 == 0u && q == 0) { z &= ~0x1Fu; keep &= ~0x1Fu; }   /* start code + NAL header */
                    const bool b0_le3 = (w.x >> 26) == 0u && (keep & 1u);
                    sm.tr[tid] = epb_group_trail(z, max(vb, 1), 0u) | (z == valid ? 0x100u : 0u);
                    __syncthreads();
                    /* zero bytes right before the group: walk back over all-zero groups */
                    u32 rin = 0;
                    const bool need = vb > 0 && (b0_le3 || (z & 1u));
                    if (need) {
                        int u = tid - 1;
                        for (;;) {
                            if (u < 0) { rin += sm.st.zrun; break; }
                            const u32 t = sm.tr[u];
                            rin += t & 0xFFu;
                            if (!(t & 0x100u)) break;
                            u--;
                        }
                    }
                    u32 ins = 0;
                    if (vb > 0 && (need || (z & (z << 1)))) {
                        u32 m3 = zero_nib(w.x & 0xFCFCFCFCu) | (zero_nib(w.y & 0xFCFCFCFCu) << 4) |
                                 (zero_nib(w.z & 0xFCFCFCFCu) << 8) | (zero_nib(w.w & 0xFCFCFCFCu) << 12);
                        ins = epb_group_mask(z, m3 & keep, vb, rin);
                    }
                    const u32 cnt = (u32)__popc(ins);
                    /* the thread holding the last valid byte carries the zero run on */
                    if (vb > 0 && vb_next == 0) {
                        u32 r2 = rin;
                        if (z == valid && !need) {                          /* all-zero group that did not need rin yet */
                            r2 = 0;
                            int u = tid - 1;
                            for (;;) {
                                if (u < 0) { r2 += sm.st.zrun; break; }
                                const u32 t = sm.tr[u];
                                r2 += t & 0xFFu;
                                if (!(t & 0x100u)) break;
                                u--;
                            }
                        }
                        sm.st.zrun_next = epb_group_trail(z, vb, r2);
                    }
                    sm.gi[tid] = (cnt << 8) | (u32)vb;
                    const int any_ins = __syncthreads_or(cnt != 0u);
                    u32 before = sm.st.ins_total, pass_ins = 0;
                    if (any_ins) before += block_excl_scan(cnt, sm.scan, pass_ins);
                    if (vb > 0) {
                        const u32 D = chunk_byte0 + 16u * (u32)q + before;          /* output byte of the group's first byte */
                        const bool clean = cnt == 0u && vb == 16;
                        bool prev_clean = tid > 0, next_clean = tid < NT - 1 && vb_next == 16;
                        if (any_ins) {
                            if (tid > 0) prev_clean = sm.gi[tid - 1] == 16u;
                            if (tid < NT - 1) next_clean = sm.gi[tid + 1] == 16u;
                        }
                        if (clean) {
                            const u32 s = (4u - (D & 3u)) & 3u;                     /* source bytes before the first aligned word */
                            if (s && !prev_clean)
                                for (u32 k = 0; k < s; k++)
                                    if (D + k < cap) fout[D + k] = (uint8_t)(w.x >> (24u - 8u * k));
                            const u32 sel = realign_selector(s);
                            u32 *dst = fout32 + ((D + s) >> 2);
                            const u32 o0 = __byte_perm(w.x, w.y, sel), o1 = __byte_perm(w.y, w.z, sel),
                                      o2 = __byte_perm(w.z, w.w, sel), o3 = __byte_perm(w.w, wn, sel);
                            if (D + s + 16u <= cap && (s == 0u || next_clean)) {
                                dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3;
                            } else {
                                if (D + s + 4u <= cap) dst[0] = o0;
                                if (D + s + 8u <= cap) dst[1] = o1;
                                if (D + s + 12u <= cap) dst[2] = o2;
                                /* the straddling word: own bytes only */
                                const u32 nb = s ? 4u - s : 4u;
                                for (u32 k = 0; k < nb; k++)
                                    if (D + s + 12u + k < cap) fout[D + s + 12u + k] = (uint8_t)(o3 >> (8u * k));
                            }
                        } else {
                            u32 pos = D;
                            const u32 ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int k = 0; k < 16; k++)
                                if (k < vb) {
                                    if ((ins >> k) & 1u) { if (pos < cap) fout[pos] = 3; pos++; }
                                    if (pos < cap) fout[pos] = (uint8_t)(ws[k >> 2] >> (24 - 8 * (k & 3)));
                                    pos++;
                                }
                        }
                    }
                    __syncthreads();
                    if (tid == 0) {
                        sm.st.ins_total += pass_ins;
                        sm.st.zrun = sm.st.zrun_next;
                    }
                }
                __syncthreads();
            }

            __syncthreads();
            if (tid == 0) {
                sm.st.bitpos = end;
                if (!last) {
                    const u32 cw = (tile0 ? 0u : w_lo) + ((w_end - (tile0 ? 0u : w_lo)) / CH) * CH;   /* base of the pass holding w_end */
                    sm.st.carry = sm.win[w_end - cw];
                }
                if (skip_spec) sm.st.last_coded = last_coded_total;
                sm.st.n_copy = 0;
                if (last) sizes[f] = total_bytes + sm.st.ins_total;
            }
            __syncthreads();
        }
    }
}

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
    int grid = n_frames < sm_count * 4 ? n_frames : sm_count * 4;
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
