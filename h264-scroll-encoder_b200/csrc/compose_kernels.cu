/*
 * compose_kernels.cu - sm_100a kernels of the P-slice composer.
 *
 * One CTA composes one frame at a time (persistent grid-stride over frames).
 * A frame is processed in tiles of TILE macroblocks; per tile:
 *
 *   load     hints of the tile + (mb_w+1) halo -> shared memory (128-bit loads)
 *   decide   per MB: A/B/C(D) neighbours from shared memory, MV predictor, P_Skip
 *            test, mvd, closed-form Exp-Golomb -> 16-byte MbEntry + bit length
 *            (replaces src/h264_writer.c:595-641, :369-453, src/bitwriter.c:50-101)
 *   skiprun  SPEC mode only: ballot masks + prefix-max -> mb_skip_run per coded MB
 *   scan     block prefix sum of bit lengths -> absolute slice bit offsets
 *   pack     output-word-owned gather: each thread builds whole 32-bit words from
 *            the MbEntries overlapping them; spliced MBs are funnel-shifted in from
 *            the pre-encoded payload (semantics of copy_bits, src/h264_writer.c:228)
 *   epb      emulation prevention (src/nal.c:24-50) in closed form: detect per word,
 *            prefix-sum the insert flags, compact; untouched words are stored directly
 *
 * The Annex-B start code, the NAL header byte and the host-written slice header
 * are treated as the first bits of the same bit string, so the shared-memory
 * image is the output image.
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

struct FrameState {
    u32 bitpos;        /* slice bits emitted so far (incl. 40-bit Annex-B prefix) */
    u32 carry;         /* partial output word at bitpos>>5 */
    int last_coded;    /* frame index of the last non-skipped MB, -1 if none */
    int zc;            /* rbsp_to_ebsp automaton state (0,1,2) after the emitted bytes */
    u32 ins_total;     /* emulation-prevention bytes inserted so far */
    u32 trailer_len;
    int last_coded_next;
    u32 carry_next;
};

struct __align__(16) ComposeSmem {
    uint2 hint[HALO_MAX + TILE];
    MbEntry ent[TILE + 1];
    u32 off[TILE + 4];
    u32 out[CH];
    unsigned short first[CH];
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
    if (warp == 0) {
        u32 w = lane < NW ? s_scan[lane] : 0u;
        u32 wi = w;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < NW) s_scan[lane] = wi - w;
        if (lane == NW - 1) s_scan[NW] = wi;
    }
    __syncthreads();
    u32 res = s_scan[warp] + inc - v;
    total = s_scan[NW];
    __syncthreads();
    return res;
}

__device__ __forceinline__ uint2 ld_hint2(const uint2 *p) { return __ldg(p); }

__global__ void __launch_bounds__(NT, 4)
compose_frames_kernel(const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                      const H264GpuMbHint *__restrict__ hints,
                      const H264GpuDynDesc *__restrict__ dyns,
                      const u32 *__restrict__ dyn_tab,
                      const uint8_t *__restrict__ payload,
                      uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    __shared__ ComposeSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        /* ---- frame setup ------------------------------------------------ */
        __syncthreads();
        if (tid < (int)(sizeof(H264GpuFrameDesc) / 4))
            ((u32 *)&sm.desc)[tid] = ((const u32 *)&descs[f])[tid];
        __syncthreads();
        const int W = sm.desc.mb_w, N = W * (int)sm.desc.mb_h;
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
                    else { uint2 s = ld_hint2(fh + g); v = make_uint4(s.x, s.y, 0u, 0u); }
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
                u32 len = 0;
                if (i < n_real) {
                    const uint2 h = sm.hint[halo + i];
                    const u32 kind = (h.y >> 8) & 0xFFu;
                    MbEntry e;
                    e.w0 = 0; e.w1 = 0; e.w2 = 0; e.w3 = 0;
                    coded = 1;
                    if (kind == H264GPU_MB_INTER) {
                        const int y = g / W, x = g - y * W;
                        const int ref = (int)(h.y & 0xFFu);
                        Nbr a = {0, 0, 0}, b = {0, 0, 0}, c = {0, 0, 0};
                        const int avA = x > 0, avB = y > 0;
                        int avC = 0;
                        if (avA) { uint2 n = sm.hint[halo + i - 1]; a = nbr_from_hint(n.x, n.y); }
                        if (avB) {
                            uint2 n = sm.hint[halo + i - W];
                            b = nbr_from_hint(n.x, n.y);
                            if (x + 1 < W) { uint2 m = sm.hint[halo + i - W + 1]; c = nbr_from_hint(m.x, m.y); avC = 1; }
                            else if (x > 0) { uint2 m = sm.hint[halo + i - W - 1]; c = nbr_from_hint(m.x, m.y); avC = 1; }
                        }
                        const int mvx = (int)(int16_t)(h.x & 0xFFFFu), mvy = (int)(int16_t)(h.x >> 16);
                        int px, py;
                        predict_mv(pred_spec, a, avA, b, avB, c, avC, ref, px, py);
                        if (skip_spec && ref == 0) {
                            int sx, sy;
                            skip_infer_mv(pred_spec, a, avA, b, avB, c, avC, sx, sy);
                            if (mvx == sx && mvy == sy) coded = 0;
                        }
                        if (coded) len = (u32)encode_p16x16_body(ref, mvx - px, mvy - py, num_refs, e);
                    } else {
                        const u32 di = h.y >> 16;
                        const u32 b0 = __ldg(tab + di), b1 = __ldg(tab + di + 1);
                        e.w1 = b0; e.w2 = b1 - b0; e.w3 = ENT_DYN;
                        len = e.w2;
                    }
                    if (!skip_spec) { e.w0 = 0x80000000u; e.w3 |= 1u; len += 1u; }   /* ue(0), :630 */
                    sm.ent[i] = e;
                }
                sm.off[i] = len;
                const u32 m = __ballot_sync(0xFFFFFFFFu, coded);
                if (lane == 0) sm.mask[i >> 5] = m;
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
                        u32 bits; int n;
                        encode_skip_prefix(run, bits, n);
                        sm.ent[i].w0 = bits;
                        sm.ent[i].w3 |= (u32)n;
                        sm.off[i] += (u32)n;
                    }
                }
                __syncthreads();
            }

            /* ---- slice trailer: trailing mb_skip_run + rbsp_stop_one_bit ---- */
            if (last && tid == 0) {
                u32 bits = 0x80000000u; int n = 1;
                if (skip_spec) {
                    const int lc = sm.st.last_coded_next;
                    const u32 run_end = (u32)(N - 1 - lc);
                    if (run_end > 0) {
                        encode_skip_prefix(run_end, bits, n);
                        bits |= 1u << (31 - n);
                        n += 1;
                    }
                }
                MbEntry e; e.w0 = bits; e.w1 = 0; e.w2 = 0; e.w3 = (u32)n;
                sm.ent[n_real] = e;
                sm.st.trailer_len = (u32)n;
            }

            /* ---- bit offsets: prefix sum of the lengths --------------------- */
            const u32 bitpos = sm.st.bitpos;
            {
                uint4 l = *(uint4 *)&sm.off[MPT * tid];
                u32 s = l.x + l.y + l.z + l.w, total;
                u32 base = bitpos + block_excl_scan(s, sm.scan, total);
                uint4 o;
                o.x = base; o.y = base + l.x; o.z = o.y + l.y; o.w = o.z + l.z;
                *(uint4 *)&sm.off[MPT * tid] = o;
                __syncthreads();
                if (tid == 0) {
                    u32 e0 = bitpos + total;
                    sm.off[n_real] = e0;                       /* sentinel / trailer start */
                    if (last) sm.off[n_real + 1] = e0 + sm.st.trailer_len;
                }
                __syncthreads();
            }
            const int n = last ? n_real + 1 : n_real;          /* entries; off[n] = end */
            const u32 end = sm.off[n];
            const u32 w_lo = bitpos >> 5;
            const u32 total_bytes = (end + 7u) >> 3;           /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* words [.., w_stop) are emitted */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end;
            const u32 cb0 = tile0 ? 0u : w_lo;
            const u32 carry_in = sm.st.carry;

            for (u32 cb = cb0; cb <= w_last; cb += CH) {
                /* markers: first[j] = smallest entry index whose start is >= bit 32j */
                for (int i = tid; i <= n; i += NT) {
                    const u32 jhi = sm.off[i] >> 5;
                    u32 jlo = (i == 0) ? cb0 : (sm.off[i - 1] >> 5) + 1u;
                    if (jlo < cb) jlo = cb;
                    const u32 jtop = min(jhi, cb + CH - 1u);
                    for (u32 j = jlo; j <= jtop; j++) sm.first[j - cb] = (unsigned short)i;
                }
                __syncthreads();

                /* pack: one 32-bit output word per thread per step */
                const u32 j_hi = min(cb + CH, w_last + 1u);
                for (u32 j = cb + tid; j < j_hi; j += NT) {
                    const u32 base = j << 5;
                    u32 acc = tile0 ? (base < 40u + 256u ? header_bits32(sm.desc.hdr_bits, nal_header, base) : 0u)
                                    : (j == w_lo ? carry_in : 0u);
                    int i = sm.first[j - cb];
                    if (i > 0) acc |= entry_contrib(sm.ent[i - 1], sm.off[i - 1], base, pay_words, n_pay_words);
                    while (i < n && sm.off[i] < base + 32u) {
                        acc |= entry_contrib(sm.ent[i], sm.off[i], base, pay_words, n_pay_words);
                        i++;
                    }
                    if (j < w_stop) sm.out[j - cb] = acc;
                    else sm.st.carry_next = acc;               /* partial word, next tile */
                }
                __syncthreads();

                /* emulation prevention + store of words [cb, e_hi) */
                const u32 e_hi = min(cb + CH, w_stop);
                const int cw = e_hi > cb ? (int)(e_hi - cb) : 0;
                const u32 chunk_byte0 = cb << 2;
                const int zc_in = sm.st.zc;
                const u32 ins_in = sm.st.ins_total;
                const int chunk_bytes = last ? (int)(min(e_hi << 2, total_bytes) - chunk_byte0) : cw * 4;
                u32 masks = 0;
                for (int it = 0; it * NT < cw; it++) {
                    const int jj = it * NT + tid;
                    if (jj < cw) {
                        const u32 j = cb + jj;
                        const u32 w = sm.out[jj];
                        bool cand = zero_byte_mask(w) != 0u;
                        if (jj > 0) cand = cand || zero_byte_mask(sm.out[jj - 1]) != 0u;
                        else cand = cand || zc_in > 0;
                        if (j == 0) cand = false;
                        if (j == 1) cand = (zero_byte_mask(w | 0xFF000000u) != 0u);
                        if (cand) {
                            int vb = min(4, chunk_bytes - 4 * jj);
                            masks |= epb_insert_mask(sm.out, jj, chunk_byte0, zc_in, vb) << (4 * it);
                        }
                    }
                }
                const int any_ins = __syncthreads_or(masks != 0u);
                if (!any_ins && ins_in == 0u) {
                    for (int jj = tid; jj < cw; jj += NT) {
                        const u32 j = cb + jj;
                        if (4u * j + 4u <= cap) fout32[j] = bswap32(sm.out[jj]);
                    }
                } else {
                    u32 running = ins_in;
                    for (int it = 0; it * NT < cw; it++) {
                        const int jj = it * NT + tid;
                        const u32 mk = (masks >> (4 * it)) & 0xFu;
                        u32 tot;
                        u32 before = running + block_excl_scan(jj < cw ? __popc(mk) : 0, sm.scan, tot);
                        if (jj < cw) {
                            const int vb = min(4, chunk_bytes - 4 * jj);
                            u32 pos = chunk_byte0 + 4u * jj + before;
                            for (int k = 0; k < vb; k++) {
                                if ((mk >> k) & 1u) { if (pos < cap) fout[pos] = 3; pos++; }
                                if (pos < cap) fout[pos] = (uint8_t)chunk_byte(sm.out, 4 * jj + k);
                                pos++;
                            }
                        }
                        running += tot;
                    }
                    if (tid == 0) sm.st.ins_total = running;
                }
                if (tid == 0 && cw > 0) sm.st.zc = epb_state_after(sm.out, chunk_bytes, chunk_byte0, zc_in);
                __syncthreads();
            }

            if (tid == 0) {
                sm.st.bitpos = end;
                sm.st.carry = sm.st.carry_next;
                if (skip_spec) sm.st.last_coded = sm.st.last_coded_next;
                if (last) sizes[f] = total_bytes + sm.st.ins_total;
            }
            __syncthreads();
        }
    }
}

/* ------------------------------------------------------------ rasteriser */

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_MAX_REGIONS = 512;

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
