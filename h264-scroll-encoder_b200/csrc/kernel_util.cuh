/*
 * kernel_util.cuh - device plumbing shared by the sm_100a kernels: warp/group scans, Ampere-style cp.async,
 * and the Hopper/Blackwell pieces the composer is built on - mbarriers (SYNCS), 1-D bulk copies through the TMA
 * unit (cp.async.bulk -> UBLKCP, L2 prefetch -> UBLKPF), bulk shared-memory clears (st.bulk -> UMEMSETS) and
 * named barriers for warp-specialised groups.
 */
#ifndef H264SCROLL_KERNEL_UTIL_CUH
#define H264SCROLL_KERNEL_UTIL_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace h264scroll {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 FULL_MASK = 0xFFFFFFFFu;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

/* ---- experiments: per-warp event trace of one CTA (COMPOSE_TRACE=1) -------------------------
   lane 0 of every warp of CTA TRACE_CTA appends (event id << 48 | clock64) to g_trace[warp][..]; TRACE(id) just before a
   barrier is the warp's arrival, just behind it the departure: arrival - previous departure = the warp's work */
#ifndef COMPOSE_TRACE
#define COMPOSE_TRACE 0
#endif
#if COMPOSE_TRACE && defined(COMPOSE_TRACE_TU)    /* (the translation unit that owns the trace buffer) */
constexpr int TRACE_EVENTS = 1024, TRACE_WARPS = 32, TRACE_CTA = 5;
__device__ unsigned long long g_trace[TRACE_WARPS][TRACE_EVENTS];
__device__ __forceinline__ void trace_event(int id, volatile unsigned int *tr_n, volatile unsigned int *tr_on) {
    if (blockIdx.x == TRACE_CTA && (threadIdx.x & 31) == 0 && *tr_on) {
        const int w = threadIdx.x >> 5;
        const unsigned int n = tr_n[w];
        if (n < (unsigned)TRACE_EVENTS) { g_trace[w][n] = ((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); tr_n[w] = n + 1; }
    }
}
#define TRACE_FIELDS unsigned int tr_n[32]; unsigned int tr_on;
#define TRACE(id) trace_event(id, sm.tr_n, &sm.tr_on)
#else
#define TRACE_FIELDS
#define TRACE(id) do { } while (0)
#endif

/* ---- cp.async (LDGSTS) ---------------------------------------------------------------------- */
__device__ __forceinline__ void cp_async16(u32 dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(u32 dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* ---- mbarriers ------------------------------------------------------------------------------ */
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
/* arrive (release at CTA scope: the thread's earlier writes are visible to whoever sees the phase complete) */
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* wait until the phase with the given parity has completed (acquire at CTA scope).  try_wait suspends the thread in
   hardware for a bounded time; a failed attempt backs off before the next one (MBAR_WAIT_NS) so that waiting warps
   stay out of the issue and fetch slots of the working ones */
#ifndef MBAR_WAIT_NS
#define MBAR_WAIT_NS 0
#endif
#ifndef MBAR_SUSPEND_HINT_NS
#define MBAR_SUSPEND_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    const u32 a = smem_u32(bar);
    u32 ok;
    for (;;) {
#if MBAR_SUSPEND_HINT_NS
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
#else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#endif
        if (ok) break;
#if MBAR_WAIT_NS
        __nanosleep(MBAR_WAIT_NS);
#endif
    }
}

/* ---- bulk copies (TMA unit, issued by ONE thread) --------------------------------------------
   global -> shared, 16-byte aligned on both sides, bytes a multiple of 16; completion is counted on the mbarrier */
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* pull [src, src + bytes) into L2 (no destination, nothing to wait for) */
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
/* zero [dst, dst + bytes) of shared memory (8-byte aligned, bytes a multiple of 8): one instruction, sm_100 */
__device__ __forceinline__ void smem_zero_bulk(void *dst_smem, u32 bytes) {
    asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(smem_u32(dst_smem)), "l"((u64)bytes) : "memory");
}

/* ---- named barriers for warp groups (id 0 is __syncthreads) --------------------------------- */
template <int ID, int COUNT>
__device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

/* ---- scans over a group of NWG warps -------------------------------------------------------- */

/*
 * Exclusive prefix sum over a group of NWG warps (gw = warp index inside the group); every thread gets its offset
 * and the total.  One barrier: consecutive scans must alternate between two s_scan buffers.
 */
template <int NWG, int BAR_ID>
__device__ __forceinline__ u32 group_excl_scan(u32 v, u32 *s_scan, int gw, u32 &total) {
    const int lane = threadIdx.x & 31;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_scan[gw] = inc;
    group_sync<BAR_ID, NWG * 32>();
    const u32 wsum = lane < NWG ? s_scan[lane] : 0u;    /* every warp scans the warp totals itself */
    u32 winc = wsum;
#pragma unroll
    for (int d = 1; d < NWG; d <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, winc, d);
        if (lane >= d) winc += t;
    }
    const u32 wexc = __shfl_sync(FULL_MASK, winc - wsum, gw);
    total = __shfl_sync(FULL_MASK, winc, NWG - 1);
    return wexc + inc - v;
}

/* exclusive prefix max (identity -1 as int); `total` = max over the group */
template <int NWG, int BAR_ID>
__device__ __noinline__ int group_excl_scan_max(int v, int *s_scan, int gw, int &total) {
    const int lane = threadIdx.x & 31;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc = max(inc, t);
    }
    if (lane == 31) s_scan[gw] = inc;
    group_sync<BAR_ID, NWG * 32>();
    const int wv = lane < NWG ? s_scan[lane] : -1;
    int winc = wv;
#pragma unroll
    for (int d = 1; d < NWG; d <<= 1) {
        int t = __shfl_up_sync(FULL_MASK, winc, d);
        if (lane >= d) winc = max(winc, t);
    }
    int wexc = __shfl_up_sync(FULL_MASK, winc, 1);
    if (lane == 0) wexc = -1;
    wexc = __shfl_sync(FULL_MASK, wexc, gw);
    total = __shfl_sync(FULL_MASK, winc, NWG - 1);
    int exc = __shfl_up_sync(FULL_MASK, inc, 1);
    if (lane == 0) exc = -1;
    return max(wexc, exc);
}

/* exclusive prefix sums of two words of packed 16-bit counters (one barrier; alternate the buffers) */
template <int NWG, int BAR_ID>
__device__ __forceinline__ void group_excl_scan_x2(u32 &lo, u32 &hi, uint2 *s_scan, int gw, u32 &tot_lo, u32 &tot_hi) {
    const int lane = threadIdx.x & 31;
    u32 il = lo, ih = hi;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 a = __shfl_up_sync(FULL_MASK, il, d), b = __shfl_up_sync(FULL_MASK, ih, d);
        if (lane >= d) { il += a; ih += b; }
    }
    if (lane == 31) s_scan[gw] = make_uint2(il, ih);
    group_sync<BAR_ID, NWG * 32>();
    const uint2 ws = lane < NWG ? s_scan[lane] : make_uint2(0u, 0u);
    u32 wl = ws.x, wh = ws.y;
#pragma unroll
    for (int d = 1; d < NWG; d <<= 1) {
        const u32 a = __shfl_up_sync(FULL_MASK, wl, d), b = __shfl_up_sync(FULL_MASK, wh, d);
        if (lane >= d) { wl += a; wh += b; }
    }
    const u32 el = __shfl_sync(FULL_MASK, wl - ws.x, gw), eh = __shfl_sync(FULL_MASK, wh - ws.y, gw);
    tot_lo = __shfl_sync(FULL_MASK, wl, NWG - 1);
    tot_hi = __shfl_sync(FULL_MASK, wh, NWG - 1);
    lo = el + il - lo;
    hi = eh + ih - hi;
}

}  /* namespace h264scroll */
#endif
