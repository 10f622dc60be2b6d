/*
 * unescape_kernels.cu - the input side of the splice on the device: EBSP -> RBSP for whole batches of NAL payloads
 * (ebsp_to_rbsp, src/nal_parser.c:67-88).
 *
 * The reference walks the payload with a zero counter and drops a byte iff
 *     zero_count >= 2  &&  byte == 0x03  &&  there is a next byte and it is <= 0x03,
 * resetting the counter at every non-zero byte (dropped or not).  Since the dropped byte itself is non-zero, the rule is
 * local: byte i goes iff b[i] == 3, b[i-1] == b[i-2] == 0 and b[i+1] <= 3 exists.  So: per 16-byte group a removal mask
 * from byte masks, a block prefix sum of the counts, a decoupled look-back over the 32 KB chunks of a payload for the
 * bytes removed before the chunk, and a compacting store (groups without removals leave as re-aligned words).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "composer_gpu.h"
#include "compose_kernels.cuh"

namespace h264scroll {

typedef uint32_t u32;
typedef uint64_t u64;

namespace {

#ifndef UNESCAPE_THREADS
#define UNESCAPE_THREADS 256
#endif
#ifndef UNESCAPE_MIN_CTAS
#define UNESCAPE_MIN_CTAS 8
#endif
constexpr int UT = UNESCAPE_THREADS;     /* threads per CTA */
constexpr int UG = 4;                    /* 16-byte groups per thread */
constexpr int UCHUNK = UT * UG * 16;     /* payload bytes per CTA */
constexpr int UW = UT / 32;
constexpr u32 UFULL = 0xFFFFFFFFu;

struct __align__(16) UnescapeSmem {
    uint4 halo_lo;                       /* the 16 bytes in front of the chunk (zero at the payload's start) */
    uint4 data[UT * UG];                 /* the chunk */
    uint4 halo_hi;                       /* the 16 bytes behind it */
    u32 gi[UT * UG];                     /* per group: removal mask << 8 | removed bytes */
    u32 gpre[UT * UG];                   /* bytes removed before a listed group */
    unsigned short list[UT * UG];        /* the candidates (the groups that go byte by byte are among them) */
    u32 n_cand;                          /* listed groups: those that may hold a removal, the partial last one, those next to the buffer's end */
    uint2 scan[UW + 1];
    u32 ticket;
    u32 base;
};

__device__ __forceinline__ u32 zero_bytes(u32 v) { return ~((((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) | 0x7F7F7F7Fu); }
/* 4-bit mask of the marked bytes (0x80 per byte) of a word loaded as it lies in memory: bit k = memory byte k */
__device__ __forceinline__ u32 nib_le(u32 m) { return (((m >> 7) & 0x01010101u) * 0x10204080u) >> 28; }
__device__ __forceinline__ u32 mask16(uint4 w, u32 x, u32 keep) {
    return nib_le(zero_bytes((w.x & keep) ^ x)) | (nib_le(zero_bytes((w.y & keep) ^ x)) << 4) |
           (nib_le(zero_bytes((w.z & keep) ^ x)) << 8) | (nib_le(zero_bytes((w.w & keep) ^ x)) << 12);
}

__device__ __forceinline__ void scan_x2(u32 &lo, u32 &hi, uint2 *s_scan, u32 &tot_lo, u32 &tot_hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 il = lo, ih = hi;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 a = __shfl_up_sync(UFULL, il, d), b = __shfl_up_sync(UFULL, ih, d);
        if (lane >= d) { il += a; ih += b; }
    }
    if (lane == 31) s_scan[warp] = make_uint2(il, ih);
    __syncthreads();
    const uint2 ws = lane < UW ? s_scan[lane] : make_uint2(0u, 0u);
    u32 wl = ws.x, wh = ws.y;
#pragma unroll
    for (int d = 1; d < UW; d <<= 1) {
        const u32 a = __shfl_up_sync(UFULL, wl, d), b = __shfl_up_sync(UFULL, wh, d);
        if (lane >= d) { wl += a; wh += b; }
    }
    const u32 el = __shfl_sync(UFULL, wl - ws.x, warp), eh = __shfl_sync(UFULL, wh - ws.y, warp);
    tot_lo = __shfl_sync(UFULL, wl, UW - 1);
    tot_hi = __shfl_sync(UFULL, wh, UW - 1);
    lo = el + il - lo;
    hi = eh + ih - hi;
}

/* decoupled look-back over the chunks of one payload: word = flag << 30 | count (1: own count, 2: inclusive sum) */
__device__ __forceinline__ u32 lookback(volatile u32 *state, int c, u32 mine) {
    if (c == 0) { state[0] = (2u << 30) | mine; return 0u; }
    state[c] = (1u << 30) | mine;
    __threadfence();
    u32 sum = 0;
    for (int p = c - 1;;) {
        const u32 v = state[p];
        if ((v >> 30) == 0u) continue;
        sum += v & 0x3FFFFFFFu;
        if ((v >> 30) == 2u || p == 0) break;
        p--;
    }
    state[c] = (2u << 30) | (sum + mine);
    return sum;
}

__global__ void __launch_bounds__(UT, UNESCAPE_MIN_CTAS)
unescape_kernel(const H264GpuUnescapeDesc *__restrict__ descs, int n_nals, int max_chunks,
                const uint8_t *__restrict__ src_base, uint8_t *__restrict__ out, u32 *__restrict__ sizes,
                u32 *__restrict__ state /* [1 + n_nals * max_chunks], zeroed */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UnescapeSmem &sm = *reinterpret_cast<UnescapeSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { sm.ticket = atomicAdd(state, 1u); sm.n_cand = 0; }   /* chunks in launch order: everything below a ticket is running or done */
    __syncthreads();
    const int ticket = (int)sm.ticket;
    /* tickets walk chunk-major over the payloads (chunk c of every payload before chunk c + 1 of any): the CTAs in flight
       then work on a few chunk indices of many payloads, so a chunk's predecessors are long done - an inclusive look-back
       word one or two steps back - instead of all being in flight with it */
    const int c = ticket / n_nals, nal = ticket - c * n_nals;
    if (c >= max_chunks) return;
    const H264GpuUnescapeDesc d = descs[nal];
    const u64 byte0 = (u64)c * UCHUNK;
    if (byte0 >= d.src_bytes) return;
    const int chunk_bytes = (int)min((u64)UCHUNK, d.src_bytes - byte0);
    const uint8_t *src = src_base + d.src_off + byte0;       /* 16-byte aligned */
    uint8_t *dst = out + d.out_off;
    /* ---- stage the chunk and 16 bytes either side ---- */
    {
        const u32 s_data = (u32)__cvta_generic_to_shared(sm.data);
        const int n16 = (chunk_bytes + 15) >> 4;             /* the last 16 bytes may reach past the payload: readable by contract */
        for (int i = tid; i < UT * UG; i += UT) {
            if (i < n16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_data + 16u * (u32)i), "l"(src + 16 * (size_t)i) : "memory");
            else sm.data[i] = make_uint4(UFULL, UFULL, UFULL, UFULL);
        }
        if (tid == 0) sm.halo_lo = byte0 ? *(const uint4 *)(src - 16) : make_uint4(UFULL, UFULL, UFULL, UFULL);
        if (tid == 32) sm.halo_hi = byte0 + (u64)chunk_bytes < d.src_bytes ? *(const uint4 *)(src + UCHUNK) : make_uint4(UFULL, UFULL, UFULL, UFULL);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    /* ---- candidates: a removal needs a 0x03 byte in the group and a zero byte in it or right in front - a cheap exact
            test that few groups pass.  Those are listed and their masks computed by as many lanes as there are candidates:
            worked out in place, one candidate would take its whole warp through the mask code (two warps in three had one) ---- */
#pragma unroll
    for (int r = 0; r < UG; r++) {
        const int g = tid + UT * r;
        const int vb = min(max(chunk_bytes - 16 * g, 0), 16);
        bool cand = false;
        if (vb > 0) {
            const uint4 w = sm.data[g];
            const u32 prev = (&sm.halo_lo)[g].w;             /* the word in front of the group (halo_lo sits right before data) */
            const u32 x3 = 0x03030303u;                      /* (the zero-byte test alone: half the instructions, twice the candidates, 0.1196 vs 0.1165 ms) */
            const u32 has3 = (((w.x ^ x3) - 0x01010101u) & ~(w.x ^ x3)) | (((w.y ^ x3) - 0x01010101u) & ~(w.y ^ x3)) |
                             (((w.z ^ x3) - 0x01010101u) & ~(w.z ^ x3)) | (((w.w ^ x3) - 0x01010101u) & ~(w.w ^ x3));
            const u32 has0 = ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) | ((w.z - 0x01010101u) & ~w.z) |
                             ((w.w - 0x01010101u) & ~w.w) | ((prev - 0x01010101u) & ~prev & 0x80000000u);
            /* (the payload's last partial group and groups next to the output buffer's end go byte by byte too: listed here,
               so that the store phase below has nothing to list) */
            cand = ((has3 & 0x80808080u) && (has0 & 0x80808080u)) || vb < 16 || byte0 + 16u * (u64)g + 20u > d.out_cap;
        }
        sm.gi[g] = 0u;
        const u32 m = __ballot_sync(UFULL, cand);
        if (m) {
            u32 at = 0;
            if (lane == 0) at = atomicAdd(&sm.n_cand, (u32)__popc(m));
            at = __shfl_sync(UFULL, at, 0);
            if (cand) sm.list[at + (u32)__popc(m & ((1u << lane) - 1u))] = (unsigned short)g;
        }
    }
    __syncthreads();
    /* ---- removal masks: one lane per candidate ---- */
    for (u32 i = (u32)tid; i < sm.n_cand; i += UT) {
        const int g = sm.list[i];
        const int vb = min(chunk_bytes - 16 * g, 16);
        const uint4 w = sm.data[g];
        const u32 prev = (&sm.halo_lo)[g].w;
        const u32 next = (&sm.data[0])[g + 1].x;             /* the word behind it (halo_hi right after data) */
        const u32 valid = vb == 16 ? 0xFFFFu : (1u << vb) - 1u;
        const u32 t = mask16(w, 0x03030303u, UFULL) & valid;
        const u32 z = mask16(w, 0u, UFULL), l3 = mask16(w, 0u, 0xFCFCFCFCu);
        const u32 zp = nib_le(zero_bytes(prev)) >> 2;            /* zero flags of the two bytes in front: bit 0 = byte -2, bit 1 = byte -1 */
        const u32 z1 = (z << 1) | (zp >> 1), z2 = (z << 2) | zp; /* byte j-1 zero, byte j-2 zero */
        /* a next byte must exist: inside the payload (bytes past its end were staged as 0xFF or are masked here) */
        const u64 left = d.src_bytes - (byte0 + 16u * (u64)g);   /* payload bytes from the group's first byte on */
        u32 nx = (l3 >> 1) | ((nib_le(zero_bytes(next & 0xFCFCFCFCu)) & 1u) << 15);
        if (left <= 16u) nx &= (1u << (u32)(left - 1u)) - 1u;   /* byte j + 1 exists only for j + 1 < left */
        const u32 rmv = t & z1 & z2 & nx;
        sm.gi[g] = (rmv << 8) | (u32)__popc(rmv);
    }
    __syncthreads();
    u32 cnt[UG];
#pragma unroll
    for (int r = 0; r < UG; r++) cnt[r] = sm.gi[tid + UT * r] & 0xFFu;
    u32 lo = cnt[0] | (cnt[1] << 16), hi = cnt[2] | (cnt[3] << 16), tlo, thi;
    scan_x2(lo, hi, sm.scan, tlo, thi);
    const u32 t0 = tlo & 0xFFFFu, t1 = tlo >> 16, t2 = thi & 0xFFFFu, t3 = thi >> 16;
    if (tid == 0) sm.base = lookback(state + 1 + (size_t)nal * max_chunks, c, t0 + t1 + t2 + t3);
    __syncthreads();
    const u32 base = sm.base;
    const u32 before[UG] = {base + (lo & 0xFFFFu), base + t0 + (lo >> 16), base + t0 + t1 + (hi & 0xFFFFu), base + t0 + t1 + t2 + (hi >> 16)};
    const u32 cap = d.out_cap;
    /* ---- compacting store ---- */
#pragma unroll
    for (int r = 0; r < UG; r++) {
        const int g = tid + UT * r;
        const int vb = min(max(chunk_bytes - 16 * g, 0), 16);
        if (vb == 0) continue;
        const u64 D64 = byte0 + 16u * (u64)g - before[r];    /* output byte of the group's first byte */
        const u32 D = (u32)D64;
        const uint4 w = sm.data[g];
        const bool roomy = byte0 + 16u * (u64)g + 20u <= cap;   /* (as the candidate test has it: removals only move the group down) */
        if (cnt[r] == 0u && vb == 16 && roomy) {
            const u32 s = (4u - (D & 3u)) & 3u;              /* bytes in front of the first aligned output word */
            u32 *o32 = (u32 *)(dst + D + s);
            if (s == 0u) {
                if ((D & 15u) == 0u) *(uint4 *)o32 = w;
                else { o32[0] = w.x; o32[1] = w.y; o32[2] = w.z; o32[3] = w.w; }
            } else {
                /* the word shared with a neighbour is stored whole only when that neighbour is a plain full group too */
                const bool prev_plain = g > 0 && sm.gi[g - 1] == 0u;
                const bool next_plain = g + 1 < UT * UG && sm.gi[g + 1] == 0u && chunk_bytes - 16 * (g + 1) >= 16;
                const u32 wn = (&sm.data[0])[g + 1].x;
                const u32 sel = 0x3210u + 0x1111u * s;
                o32[0] = __byte_perm(w.x, w.y, sel);
                o32[1] = __byte_perm(w.y, w.z, sel);
                o32[2] = __byte_perm(w.z, w.w, sel);
                const u32 o3 = __byte_perm(w.w, wn, sel);
                if (next_plain) o32[3] = o3;
                else {                                       /* (1..3 bytes, no loop: the kernel is issue bound) */
                    uint8_t *t = dst + D + s + 12u;
                    t[0] = (uint8_t)o3;
                    if (s < 3u) t[1] = (uint8_t)(o3 >> 8);
                    if (s < 2u) t[2] = (uint8_t)(o3 >> 16);
                }
                if (!prev_plain) {
                    uint8_t *h = dst + D;
                    h[0] = (uint8_t)w.x;
                    if (s > 1u) h[1] = (uint8_t)(w.x >> 8);
                    if (s > 2u) h[2] = (uint8_t)(w.x >> 16);
                }
            }
        } else {
            /* a removal inside, the payload's last partial group, or the end of the output buffer in reach: byte by byte,
               but not here - one such group would hold its whole warp up.  They are among the candidates: dealt out below */
            sm.gpre[g] = before[r];
        }
    }
    __syncthreads();
    for (u32 i = (u32)tid; i < 4u * sm.n_cand; i += UT) {   /* four lanes per group, a word each: the phase is as long as its longest lane */
        const int g = sm.list[i >> 2];
        const u32 sub = i & 3u;
        const int vbg = min(chunk_bytes - 16 * g, 16);
        if ((sm.gi[g] & 0xFFu) == 0u && vbg == 16 && byte0 + 16u * (u64)g + 20u <= cap) continue;   /* turned out plain: stored above */
        const int vbs = vbg - 4 * (int)sub;                  /* valid bytes from this word on */
        if (vbs <= 0) continue;
        const u32 wv = ((const u32 *)sm.data)[4 * g + (int)sub];
        const u32 rmg = sm.gi[g] >> 8;
        u64 pos = byte0 + 16u * (u64)g - sm.gpre[g] + 4u * sub - (u32)__popc(rmg & ((1u << (4u * sub)) - 1u));
        u32 rs = rmg >> (4u * sub);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < vbs) {
                if (!(rs & 1u)) { if (pos < cap) dst[pos] = (uint8_t)(wv >> (8 * k)); pos++; }
                rs >>= 1;
            }
        }
    }
    if (tid == 0 && byte0 + (u64)chunk_bytes >= d.src_bytes) sizes[nal] = (u32)(d.src_bytes - (base + t0 + t1 + t2 + t3));
    (void)lane;
}


/* ------------------------------------------------------------ Annex-B scan
 *
 * nal_parser_next (src/nal_parser.c:14-65) for a whole stream: a NAL unit starts behind every 00 00 01 (a 4-byte start
 * code is the same pattern one byte later, so both give the same position) and ends before the next start code, minus
 * trailing zero bytes.  Pass 1 finds the start positions in stream order (a CTA counts the patterns in its 16 KB,
 * looks back for the number before it, writes its positions); pass 2 turns them into {offset, size, header} records.
 */
constexpr int PT = 256;                  /* threads per CTA: 16 bytes each... */
constexpr int PG = 4;                    /* ...times 4 */
constexpr int PCHUNK = PT * PG * 16;

__global__ void __launch_bounds__(PT)
annexb_starts_kernel(const uint8_t *__restrict__ data, unsigned long long n_bytes, u32 *__restrict__ starts, u32 max_nals,
                     u32 *__restrict__ state /* [1 + chunks]: ticket, look-back words; zeroed */) {
    __shared__ u32 s_ticket, s_base;
    __shared__ u32 s_warp[PT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(state, 1u);
    __syncthreads();
    const int c = (int)s_ticket;
    const u64 byte0 = (u64)c * PCHUNK;
    if (byte0 >= n_bytes) return;
    /* a thread owns PG consecutive 16-byte groups: pattern starts i in [g0, g0 + 16 PG), bytes i, i+1, i+2 */
    const u64 t0 = byte0 + (u64)tid * (16 * PG);
    u32 found[PG];
    u32 cnt = 0;
#pragma unroll
    for (int r = 0; r < PG; r++) {
        const u64 g0 = t0 + 16u * (u64)r;
        found[r] = 0u;
        if (g0 < n_bytes) {
            const uint4 w = *(const uint4 *)(data + g0);      /* the buffer is readable 16 bytes past its end */
            const u32 nx = *(const u32 *)(data + g0 + 16);
            const u32 z = mask16(w, 0u, UFULL) | ((nib_le(zero_bytes(nx)) & 3u) << 16);           /* bytes 0..17 == 0 */
            const u32 o = mask16(w, 0x01010101u, UFULL) | ((nib_le(zero_bytes(nx ^ 0x01010101u)) & 3u) << 16);   /* == 1 */
            u32 m = z & (z >> 1) & (o >> 2) & 0xFFFFu;        /* bit i: bytes i, i+1 zero and byte i+2 one */
            const u64 left = n_bytes - g0;                    /* the pattern needs i + 2 < left */
            if (left < 18u) m &= left > 2u ? (1u << (u32)(left - 2u)) - 1u : 0u;
            found[r] = m;
            cnt += (u32)__popc(m);
        }
    }
    /* block exclusive scan of the counts (stream order = thread order) */
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(UFULL, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    u32 wbase = 0, total = 0;
    for (int w = 0; w < PT / 32; w++) { if (w < warp) wbase += s_warp[w]; total += s_warp[w]; }
    if (tid == 0) s_base = lookback(state + 1, c, total);
    __syncthreads();
    u32 k = s_base + wbase + inc - cnt;
#pragma unroll
    for (int r = 0; r < PG; r++) {
        u32 m = found[r];
        while (m) {
            const int i = __ffs((int)m) - 1;
            m &= m - 1u;
            /* max_nals + 1 positions are kept: record max_nals - 1 ends at the start code behind it (slot max_nals holds the
               count, so the extra position lives one word further) */
            if (k < max_nals) starts[k] = (u32)(t0 + 16u * (u64)r + (u64)i + 3u);
            else if (k == max_nals) starts[max_nals + 1u] = (u32)(t0 + 16u * (u64)r + (u64)i + 3u);
            k++;
        }
    }
    if (tid == PT - 1 && byte0 + PCHUNK >= n_bytes) starts[max_nals] = k;      /* the last chunk knows the count */
}

__global__ void __launch_bounds__(256)
annexb_units_kernel(const uint8_t *__restrict__ data, unsigned long long n_bytes, const u32 *__restrict__ starts, u32 max_nals,
                    H264GpuNalUnit *__restrict__ units) {
    const u32 n_all = starts[max_nals], n = min(n_all, max_nals);
    for (u32 k = blockIdx.x * 256u + threadIdx.x; k < n; k += gridDim.x * 256u) {
        const u64 b = starts[k];
        /* the unit ends in front of the next start code - also when that one opens a unit beyond max_nals */
        u64 e = k + 1u < n_all ? (u64)starts[k + 1u < max_nals ? k + 1u : max_nals + 1u] - 3u : n_bytes;
        while (e > b && data[e - 1u] == 0) e--;              /* "remove trailing zeros before next start code" */
        H264GpuNalUnit u;
        u.offset = (u32)b + 1u;                              /* payload behind the header byte, as NALUnit.data */
        u.size = e > b ? (u32)(e - b - 1u) : 0u;
        u.nal_ref_idc = e > b ? (data[b] >> 5) & 3u : 0u;
        u.nal_unit_type = e > b ? data[b] & 31u : 0u;
        units[k] = u;
    }
}

}  // namespace

cudaError_t launch_unescape(const H264GpuUnescapeDesc *d_desc, int n_nals, int max_chunks, const uint8_t *d_src,
                            uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_state, cudaStream_t stream) {
    if (n_nals <= 0 || max_chunks <= 0) return cudaSuccess;
    const int smem = (int)sizeof(UnescapeSmem);
    /* (per call: function attributes belong to the calling thread's current device) */
    cudaError_t e = cudaFuncSetAttribute(unescape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)n_nals * (unsigned)max_chunks;
    e = cudaMemsetAsync(d_state, 0, sizeof(uint32_t) * (1 + (size_t)grid), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_sizes, 0, sizeof(uint32_t) * (size_t)n_nals, stream);   /* an empty payload (or one with more chunks than max_chunks) reports 0 */
    if (e != cudaSuccess) return e;
    unescape_kernel<<<grid, UT, smem, stream>>>(d_desc, n_nals, max_chunks, d_src, d_out, d_sizes, d_state);
    return cudaGetLastError();
}

int unescape_chunk_bytes() { return UCHUNK; }

cudaError_t launch_parse_annexb(const uint8_t *d_data, size_t n_bytes, H264GpuNalUnit *d_units, uint32_t max_nals,
                                uint32_t *d_scratch /* 3 + max_nals + chunks */, cudaStream_t stream) {
    if (n_bytes == 0 || max_nals == 0) return cudaSuccess;
    const unsigned chunks = (unsigned)((n_bytes + PCHUNK - 1) / PCHUNK);
    uint32_t *starts = d_scratch, *state = d_scratch + max_nals + 2;    /* starts[0 .. max_nals), the count, one more start */
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, sizeof(uint32_t) * ((size_t)max_nals + 3 + chunks), stream);
    if (e != cudaSuccess) return e;
    annexb_starts_kernel<<<chunks, PT, 0, stream>>>(d_data, (unsigned long long)n_bytes, starts, max_nals, state);
    unsigned g2 = (max_nals + 255u) / 256u;
    if (g2 > 148u * 8u) g2 = 148u * 8u;
    annexb_units_kernel<<<g2, 256, 0, stream>>>(d_data, (unsigned long long)n_bytes, starts, max_nals, d_units);
    return cudaGetLastError();
}

size_t parse_annexb_scratch_words(size_t n_bytes, uint32_t max_nals) { return (size_t)max_nals + 3 + (n_bytes + PCHUNK - 1) / PCHUNK; }

}  /* namespace h264scroll */
