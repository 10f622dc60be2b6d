/*
 * compose_kernels.cu - sm_100a kernels of the P-slice composer.
 *
 * One CTA composes one frame at a time (persistent grid-stride over frames).
 * A frame is processed in tiles of TILE macroblocks; per tile:
 *
 *   load     hints of the tile + (mb_w+1) halo -> shared memory (128-bit loads)
 *   decide   per MB: A/B/C(D) neighbours from shared memory, MV predictor, P_Skip
 *            test, mvd, closed-form Exp-Golomb -> MbEntry + bit length
 *            (replaces src/h264_writer.c:595-641, :369-453, src/bitwriter.c:50-101);
 *            a uniform neighbourhood (A = B = C = own hint) short-cuts to mvd = 0
 *   skiprun  SPEC mode only: ballot masks + prefix-max -> mb_skip_run per coded MB
 *   scan     block prefix sum of bit lengths -> absolute slice bit offsets
 *   pack     output words in which a macroblock starts are gathered by an owner
 *            thread from the entries overlapping them; words lying wholly inside a
 *            spliced macroblock are funnel-shifted in from the pre-encoded payload by
 *            a whole warp with coalesced loads (semantics of copy_bits,
 *            src/h264_writer.c:228)
 *   epb      emulation prevention (src/nal.c:24-50) in closed form: zero-byte
 *            detect, prefix-max of the last non-zero byte, prefix-sum of the insert
 *            flags, compact; warps without insertions store re-aligned 32-bit words
 *
 * The Annex-B start code, the NAL header byte and the host-written slice header
 * are the first bits of the same bit string, so the shared-memory image is the
 * output image.
 */
#include <cuda_runtime.h>
#include "mb_codec.cuh"
#include "compose_kernels.cuh"

namespace h264scroll {

constexpr int NT = COMPOSE_THREADS;      /* threads per CTA */
constexpr int TILE = 1024;               /* macroblocks per tile */
constexpr int MPT = TILE / NT;           /* macroblocks per thread per tile */
constexpr int HALO_MAX = COMPOSE_MAX_MB_W + 1;
constexpr int CH = 1024;                 /* output words per pack chunk */
constexpr int NW = NT / 32;
constexpr int WPT = CH / NT;             /* output words per thread per chunk */

struct FrameState {
    u32 bitpos;        /* slice bits emitted so far (incl. 40-bit Annex-B prefix) */
    u32 carry;         /* partial output word at bitpos>>5 */
    int last_coded;    /* frame index of the last non-skipped MB, -1 if none */
    int zc;            /* rbsp_to_ebsp automaton state (0,1,2) after the emitted bytes */
    u32 ins_total;     /* emulation-prevention bytes inserted so far */
    u32 trailer_len;
    int last_coded_next;
    u32 carry_next;
    u32 n_long;        /* entries of the current tile that may own interior words (spliced / > 32 bits) */
    int zc_next;
};

struct __align__(16) ComposeSmem {
    uint2 hint[HALO_MAX + TILE];
    MbEntry ent[TILE + 1];
    u32 off[TILE + 4];                   /* lengths, then offsets; entry kind in the top two bits */
    u32 out[CH];
    u32 first[CH];                       /* (chunk tag << 16) | first entry starting in the word */
    unsigned short longlist[TILE];
    u32 mask[TILE / 32];
    int prev[TILE / 32];
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
        u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    const u32 wsum = lane < NW ? s_scan[lane] : 0u;    /* every warp scans the NW warp totals itself */
    u32 winc = wsum;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
        if (lane >= d) winc += t;
    }
    const u32 wexc = __shfl_sync(0xFFFFFFFFu, winc - wsum, warp);
    total = __shfl_sync(0xFFFFFFFFu, winc, NW - 1);
    __syncthreads();                                   /* s_scan may be reused by the caller */
    return wexc + inc - v;
}

/* exclusive prefix max over the CTA (identity `ident`); `total` = max over all threads */
__device__ __forceinline__ int block_excl_scan_max(int v, int ident, int *s_scan, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc = max(inc, t);
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    const int wv = lane < NW ? s_scan[lane] : ident;
    int winc = wv;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
        if (lane >= d) winc = max(winc, t);
    }
    int wexc = __shfl_up_sync(0xFFFFFFFFu, winc, 1);
    if (lane == 0) wexc = ident;
    wexc = __shfl_sync(0xFFFFFFFFu, wexc, warp);
    total = __shfl_sync(0xFFFFFFFFu, winc, NW - 1);
    __syncthreads();
    int exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) exc = ident;
    return max(wexc, exc);
}

__global__ void __launch_bounds__(NT, 4)
compose_frames_kernel(const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                      const H264GpuMbHint *__restrict__ hints,
                      const H264GpuDynDesc *__restrict__ dyns,
                      const u32 *__restrict__ dyn_tab,
                      const uint8_t *__restrict__ payload,
                      uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    __shared__ ComposeSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 chunk_tag = 0;                    /* bumps per pack chunk; stale first[] entries never match */

    for (int jj = tid; jj < CH; jj += NT) sm.first[jj] = 0xFFFFFFFFu;

    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        /* ---- frame setup ------------------------------------------------ */
        __syncthreads();
        if (tid < (int)(sizeof(H264GpuFrameDesc) / 4))
            ((u32 *)&sm.desc)[tid] = ((const u32 *)&descs[f])[tid];
        __syncthreads();
        const int W = sm.desc.mb_w, N = W * (int)sm.desc.mb_h;
        const u32 w_magic = 0xFFFFFFFFu / (u32)W + 1u;         /* g / W == umulhi(g, magic): exact for g < 2^20, W < 2^10 */
        const int num_refs = sm.desc.num_refs;
        const int skip_spec = (sm.desc.flags & H264GPU_SKIP_SPEC) != 0;
        const int pred_spec = (sm.desc.flags & H264GPU_PRED_SPEC) != 0;
        const u32 cap = sm.desc.out_cap;
        const u32 nal_header = sm.desc.nal_header;
        const u32 dyn_id = sm.desc.dyn_id;
        if (dyn_id != H264GPU_NO_DYN && tid < (int)(sizeof(H264GpuDynDesc) / 4))
            ((u32 *)&sm.dyn)[tid] = ((const u32 *)&dyns[dyn_id])[tid];
        if (tid == 0) {
            sm.st.bitpos = 40u + sm.desc.hdr_nbits;
            sm.st.carry = 0;
            sm.st.last_coded = -1;
            sm.st.zc = 0;
            sm.st.ins_total = 0;
            sm.st.n_long = 0;
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
            const int n_real = min(TILE, N - t0);
            const bool last = (t0 + TILE >= N);
            const bool tile0 = (t0 == 0);
            const int halo = min(t0, W + 1);

            /* ---- load hints: tile + halo (hint_off even => 16-byte aligned pairs) */
            {
                const int first_g = t0 - halo;            /* first hint index needed */
                const int cnt = halo + n_real;
                const int g_al = first_g & ~1;             /* align down to a pair */
                const int lead = first_g - g_al;           /* 0 or 1 */
                const int pairs = (cnt + lead + 1) >> 1;
                const uint4 *src = (const uint4 *)(fh + g_al);
                for (int p = tid; p < pairs; p += NT) {
                    int g = g_al + 2 * p;
                    uint4 v;
                    if (g + 1 < N) v = __ldg(src + p);
                    else { uint2 s2 = __ldg(fh + g); v = make_uint4(s2.x, s2.y, 0u, 0u); }
                    int idx = 2 * p - lead;
                    if (idx >= 0) sm.hint[idx] = make_uint2(v.x, v.y);
                    if (idx + 1 < cnt) sm.hint[idx + 1] = make_uint2(v.z, v.w);
                }
            }
            __syncthreads();

            /* ---- decide + code ------------------------------------------- */
#pragma unroll
            for (int k = 0; k < MPT; k++) {
                const int i = k * NT + tid;
                const int g = t0 + i;
                int coded = 0;
                u32 len = 0, kind = KIND_SHORT;
                if (i < n_real) {
                    const uint2 h = sm.hint[halo + i];
                    const u32 hk = (h.y >> 8) & 0xFFu;
                    MbEntry e;
                    e.w0 = 0; e.w1 = 0; e.w2 = 0; e.w3 = 0;
                    coded = 1;
                    if (hk == H264GPU_MB_INTER) {
                        const int y = W == 1 ? g : (int)__umulhi((u32)g, w_magic), x = g - y * W;
                        const int ref = (int)(h.y & 0xFFu);
                        const int avA = x > 0, avB = y > 0;
                        const int avC = avB && (x + 1 < W || x > 0);
                        uint2 ha = make_uint2(0u, 0u), hb = ha, hc = ha;
                        if (avA) ha = sm.hint[halo + i - 1];
                        if (avB) hb = sm.hint[halo + i - W];
                        if (avC) hc = sm.hint[halo + i - W + (x + 1 < W ? 1 : -1)];
                        const u32 own_lo = h.x, own_hi = h.y & 0xFFFFu;
                        const bool uniform = avA && avB && avC &&
                                             ha.x == own_lo && hb.x == own_lo && hc.x == own_lo &&
                                             (ha.y & 0xFFFFu) == own_hi && (hb.y & 0xFFFFu) == own_hi &&
                                             (hc.y & 0xFFFFu) == own_hi;
                        int blen;
                        if (uniform) {
                            /* A = B = C = this MB: every predictor rule yields the MB's own MV -> mvd 0;
                             * with ref 0 that is also exactly the inferred P_Skip motion */
                            if (skip_spec && ref == 0) coded = 0;
                            blen = encode_p16x16_body(ref, 0, 0, num_refs, e);
                        } else {
                            const Nbr a = nbr_from_hint(ha.x, ha.y), b = nbr_from_hint(hb.x, hb.y),
                                      c = nbr_from_hint(hc.x, hc.y);
                            const int mvx = (int)(int16_t)(h.x & 0xFFFFu), mvy = (int)(int16_t)(h.x >> 16);
                            int px, py;
                            predict_mv(pred_spec, a, avA, b, avB, c, avC, ref, px, py);
                            if (skip_spec && ref == 0) {
                                int sx, sy;
                                skip_infer_mv(pred_spec, a, avA, b, avB, c, avC, sx, sy);
                                if (mvx == sx && mvy == sy) coded = 0;
                            }
                            blen = encode_p16x16_body(ref, mvx - px, mvy - py, num_refs, e);
                        }
                        if (coded) {
                            len = (u32)blen;
                            if (!skip_spec) { kind = finish_inline_entry(e, 0x80000000u, 1, blen); len += 1u; }  /* ue(0), :630 */
                        }
                    } else {
                        const u32 di = h.y >> 16;
                        const u32 b0 = __ldg(tab + di), b1 = __ldg(tab + di + 1);
                        e.w1 = b0; e.w2 = b1 - b0; e.w3 = ENT_DYN;
                        len = e.w2;
                        kind = KIND_DYN;
                        if (!skip_spec) { e.w0 = 0x80000000u; e.w3 |= 1u; len += 1u; }
                    }
                    sm.ent[i] = e;
                }
                sm.off[i] = len | (kind << OFF_KIND_SHIFT);
                const u32 m = __ballot_sync(0xFFFFFFFFu, coded);
                if (lane == 0) sm.mask[i >> 5] = m;
                /* entries that can own whole output words are listed for the cooperative copy */
                const bool listed = kind != KIND_SHORT;
                const u32 lm = __ballot_sync(0xFFFFFFFFu, listed);
                if (lm) {
                    u32 base = 0;
                    if (lane == 0) base = atomicAdd(&sm.st.n_long, (u32)__popc(lm));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (listed) sm.longlist[base + __popc(lm & ((1u << lane) - 1u))] = (unsigned short)i;
                }
            }
            __syncthreads();

            /* ---- mb_skip_run (SPEC mode) ----------------------------------- */
            if (skip_spec) {
                if (warp == 0) {
                    const u32 m = sm.mask[lane];
                    int v = m ? (t0 + 32 * lane + 31 - __clz((int)m)) : -1;
                    int inc = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        int t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                        if (lane >= d) inc = max(inc, t);
                    }
                    int exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
                    if (lane == 0) exc = -1;
                    sm.prev[lane] = max(exc, sm.st.last_coded);
                    if (lane == 31) sm.st.last_coded_next = max(inc, sm.st.last_coded);
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < MPT; k++) {
                    const int i = k * NT + tid;
                    const u32 m = sm.mask[i >> 5];
                    if ((m >> (i & 31)) & 1u) {
                        const u32 below = m & ((1u << (i & 31)) - 1u);
                        const int prev = below ? (t0 + (i & ~31) + 31 - __clz((int)below)) : sm.prev[i >> 5];
                        const u32 run = (u32)(t0 + i - 1 - prev);
                        u32 bits; int pn;
                        encode_skip_prefix(run, bits, pn);
                        const u32 lk = sm.off[i];
                        u32 kind = lk >> OFF_KIND_SHIFT;
                        const u32 blen = lk & OFF_MASK;
                        if (kind == KIND_DYN) {
                            sm.ent[i].w0 = bits;
                            sm.ent[i].w3 |= (u32)pn;
                        } else {
                            MbEntry e = sm.ent[i];
                            kind = finish_inline_entry(e, bits, pn, (int)blen);
                            sm.ent[i] = e;
                            if (kind == KIND_LONG) sm.longlist[atomicAdd(&sm.st.n_long, 1u)] = (unsigned short)i;
                        }
                        sm.off[i] = (blen + (u32)pn) | (kind << OFF_KIND_SHIFT);
                    }
                }
                __syncthreads();
            }

            /* ---- slice trailer: trailing mb_skip_run + rbsp_stop_one_bit ---- */
            if (last && tid == 0) {
                u32 bits = 0x80000000u; int pn = 1;
                if (skip_spec) {
                    const int lc = sm.st.last_coded_next;
                    const u32 run_end = (u32)(N - 1 - lc);
                    if (run_end > 0) {
                        encode_skip_prefix(run_end, bits, pn);
                        bits |= 1u << (31 - pn);
                        pn += 1;
                    }
                }
                MbEntry e; e.w0 = bits; e.w1 = 0; e.w2 = 0; e.w3 = 0;
                sm.ent[n_real] = e;                            /* a KIND_SHORT entry */
                sm.st.trailer_len = (u32)pn;
            }

            /* ---- bit offsets: prefix sum of the lengths --------------------- */
            const u32 bitpos = sm.st.bitpos;
            {
                uint4 l = *(uint4 *)&sm.off[MPT * tid];
                const u32 k0 = l.x & ~OFF_MASK, k1 = l.y & ~OFF_MASK, k2 = l.z & ~OFF_MASK, k3 = l.w & ~OFF_MASK;
                l.x &= OFF_MASK; l.y &= OFF_MASK; l.z &= OFF_MASK; l.w &= OFF_MASK;
                u32 sum4 = l.x + l.y + l.z + l.w, total;
                u32 base = bitpos + block_excl_scan(sum4, sm.scan, total);
                uint4 o;
                o.x = base; o.y = base + l.x; o.z = o.y + l.y; o.w = o.z + l.z;
                o.x |= k0; o.y |= k1; o.z |= k2; o.w |= k3;
                *(uint4 *)&sm.off[MPT * tid] = o;
                __syncthreads();
                if (tid == 0) {
                    u32 e0 = bitpos + total;
                    sm.off[n_real] = e0;                       /* sentinel / trailer start (KIND_SHORT) */
                    if (last) sm.off[n_real + 1] = e0 + sm.st.trailer_len;
                }
                __syncthreads();
            }
            const int n = last ? n_real + 1 : n_real;          /* entries; off[n] = end */
            const u32 end = sm.off[n] & OFF_MASK;
            const u32 w_lo = bitpos >> 5;
            const u32 total_bytes = (end + 7u) >> 3;           /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* words [.., w_stop) are emitted */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end;
            const u32 cb0 = tile0 ? 0u : w_lo;
            const u32 carry_in = sm.st.carry;
            const u32 n_long = sm.st.n_long;

            for (u32 cb = cb0; cb <= w_last; cb += CH) {
                chunk_tag = (chunk_tag + 1u) & 0xFFFFu;
                if (chunk_tag == 0u) {                         /* tag wrapped: forget every stale marker */
                    for (int jj = tid; jj < CH; jj += NT) sm.first[jj] = 0xFFFFFFFFu;
                    chunk_tag = 1u;
                    __syncthreads();
                }
                const u32 tagbits = chunk_tag << 16;

                /* markers: for every word in which an entry starts, the first such entry */
                for (int i = tid; i <= n; i += NT) {
                    const u32 j = (sm.off[i] & OFF_MASK) >> 5;
                    const bool first_in_word = (i == 0) || (((sm.off[i - 1] & OFF_MASK) >> 5) != j);
                    if (first_in_word && j >= cb && j < cb + CH) sm.first[j - cb] = tagbits | (u32)i;
                }
                if (tile0 && cb + (u32)tid < w_lo)             /* words holding only header bits */
                    sm.first[tid] = tagbits;
                __syncthreads();

                /* pack (a): words in which an entry starts -> gathered by one thread each */
                const u32 j_hi = min(cb + CH, w_last + 1u);
                for (u32 j = cb + tid; j < j_hi; j += NT) {
                    const u32 fm = sm.first[j - cb];
                    if ((fm & 0xFFFF0000u) != tagbits) continue;           /* interior word, see (b) */
                    const u32 base = j << 5;
                    u32 acc = tile0 ? (base < 40u + 256u ? header_bits32(sm.desc.hdr_bits, nal_header, base) : 0u)
                                    : (j == w_lo ? carry_in : 0u);
                    int i = (int)(fm & 0xFFFFu);
                    if (i > 0) {
                        const u32 o = sm.off[i - 1];
                        if ((o >> OFF_KIND_SHIFT) == KIND_SHORT) acc |= short_contrib(sm.ent[i - 1].w0, o & OFF_MASK, base);
                        else acc |= entry_contrib(sm.ent[i - 1], o & OFF_MASK, base, pay_words, n_pay_words);
                    }
                    while (i < n) {
                        const u32 o = sm.off[i];
                        if ((o & OFF_MASK) >= base + 32u) break;
                        if ((o >> OFF_KIND_SHIFT) == KIND_SHORT) acc |= short_contrib(sm.ent[i].w0, o & OFF_MASK, base);
                        else acc |= entry_contrib(sm.ent[i], o & OFF_MASK, base, pay_words, n_pay_words);
                        i++;
                    }
                    if (j < w_stop) sm.out[j - cb] = acc;
                    else sm.st.carry_next = acc;               /* partial word, carried into the next tile */
                }
                /* pack (b): words lying wholly inside one long entry -> a warp per entry, a lane per word */
                for (u32 d = warp; d < n_long; d += NW) {
                    const int i = sm.longlist[d];
                    const u32 o = sm.off[i] & OFF_MASK;
                    const u32 o_next = sm.off[i + 1] & OFF_MASK;
                    u32 ja = (o >> 5) + 1u, jb = o_next >> 5;              /* interior words [ja, jb) */
                    if (ja < cb) ja = cb;
                    if (jb > cb + CH) jb = cb + CH;
                    if (ja >= jb) continue;
                    const MbEntry e = sm.ent[i];
                    for (u32 j = ja + lane; j < jb; j += 32)
                        sm.out[j - cb] = entry_contrib(e, o, j << 5, pay_words, n_pay_words);
                }
                __syncthreads();

                /* ---- emulation prevention + store of words [cb, e_hi) ------ */
                const u32 e_hi = min(cb + CH, w_stop);
                const int cw = e_hi > cb ? (int)(e_hi - cb) : 0;
                const u32 chunk_byte0 = cb << 2;
                const int zc_in = sm.st.zc;
                const u32 ins_in = sm.st.ins_total;
                const int chunk_bytes = last ? (int)(min(e_hi << 2, total_bytes) - chunk_byte0) : cw * 4;
                u32 wv[WPT];
                bool cand = false;
#pragma unroll
                for (int it = 0; it < WPT; it++) {
                    const int jj = it * NT + tid;
                    wv[it] = 0xFFFFFFFFu;
                    if (jj < cw) {
                        const u32 w = sm.out[jj];
                        wv[it] = w;
                        const u32 j = cb + jj;
                        /* protection needs two zero bytes in a row somewhere in the chunk */
                        bool c = zero_byte_mask(j == 1u ? (w | 0xFF000000u) : w) != 0u;
                        if (j == 0u) c = false;
                        cand = cand || c;
                    }
                }
                if (tid == 0 && zc_in > 0) cand = true;
                const int any_cand = __syncthreads_or(cand);
                u32 masks = 0;
                int any_ins = 0;
                if (any_cand) {
                    int running_nz = -1 - zc_in;                           /* carried automaton state */
#pragma unroll
                    for (int it = 0; it < WPT; it++) {
                        if (it * NT < cw) {
                            const int jj = it * NT + tid;
                            const int vb = jj < cw ? min(4, chunk_bytes - 4 * jj) : 0;
                            const int lnz = word_last_nonzero(wv[it], vb, 4 * jj);
                            int tot;
                            int prev_nz = block_excl_scan_max(lnz, EPB_NONE, (int *)sm.scan, tot);
                            prev_nz = max(prev_nz, running_nz);
                            if (vb > 0) masks |= epb_word_mask(wv[it], vb, 4 * jj, chunk_byte0, prev_nz) << (4 * it);
                            running_nz = max(running_nz, tot);
                        }
                    }
                    any_ins = __syncthreads_or(masks != 0u);
                    if (tid == 0) {
                        const int r_end = chunk_bytes - 1 - running_nz;    /* zero bytes closing the chunk */
                        sm.st.zc_next = r_end == 0 ? 0 : ((r_end & 1) ? 1 : 2);
                    }
                } else if (tid == 0) sm.st.zc_next = 0;

                u32 running = ins_in;
#pragma unroll
                for (int it = 0; it < WPT; it++) {
                    if (it * NT < cw) {
                        const int jj = it * NT + tid;
                        const u32 mk = (masks >> (4 * it)) & 0xFu;
                        u32 before = running;
                        if (any_ins) {
                            u32 tot;
                            before += block_excl_scan(jj < cw ? (u32)__popc(mk) : 0u, sm.scan, tot);
                            running += tot;
                        }
                        const int vb = jj < cw ? min(4, chunk_bytes - 4 * jj) : 0;
                        const u32 w = wv[it];
                        /* a warp covers 32 consecutive words: without insertions (and no partial tail word)
                         * they move as re-aligned 32-bit words, otherwise byte by byte */
                        const bool byte_path = __any_sync(0xFFFFFFFFu, vb > 0 && (mk != 0u || vb < 4));
                        const u32 prevw = __shfl_up_sync(0xFFFFFFFFu, w, 1);
                        const u32 nextvb = __shfl_down_sync(0xFFFFFFFFu, (u32)vb, 1);
                        if (vb > 0) {
                            const u32 src_byte = chunk_byte0 + 4u * (u32)jj;
                            if (byte_path) {
                                u32 pos = src_byte + before;
                                for (int k = 0; k < vb; k++) {
                                    if ((mk >> k) & 1u) { if (pos < cap) fout[pos] = 3; pos++; }
                                    if (pos < cap) fout[pos] = (uint8_t)(w >> (24 - 8 * k));
                                    pos++;
                                }
                            } else {
                                const u32 sft = before & 3u;
                                if (sft == 0u) {
                                    const u32 pos = src_byte + before;
                                    if (pos + 4u <= cap) fout32[pos >> 2] = bswap32(w);
                                } else {
                                    const bool head = (lane == 0);
                                    const bool tail = (lane == 31) || (nextvb == 0u);
                                    if (!head) {
                                        /* out word covering source bytes [src_byte - sft, src_byte - sft + 4) */
                                        const u32 v = (prevw << (8u * (4u - sft))) | (w >> (8u * sft));
                                        const u32 pos = src_byte - sft + before;
                                        if (pos + 4u <= cap) fout32[pos >> 2] = bswap32(v);
                                    } else {
                                        u32 pos = src_byte + before;
                                        for (u32 k = 0; k < 4u - sft; k++, pos++)
                                            if (pos < cap) fout[pos] = (uint8_t)(w >> (24 - 8 * k));
                                    }
                                    if (tail) {
                                        u32 pos = src_byte + 4u - sft + before;
                                        for (u32 k = 4u - sft; k < 4u; k++, pos++)
                                            if (pos < cap) fout[pos] = (uint8_t)(w >> (24 - 8 * k));
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    sm.st.ins_total = running;
                    if (cw > 0) sm.st.zc = sm.st.zc_next;
                }
                __syncthreads();
            }

            if (tid == 0) {
                sm.st.bitpos = end;
                sm.st.carry = sm.st.carry_next;
                if (skip_spec) sm.st.last_coded = sm.st.last_coded_next;
                sm.st.n_long = 0;
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
