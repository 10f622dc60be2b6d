/*
 * compose_kernels.cu - sm_100a kernels of the P-slice composer.
 *
 * compose_frames_kernel (v7): a persistent CTA composes one frame after the other (grid-stride over frames) with two
 * warp-specialised groups that run as a pipeline over double-buffered slice windows in shared memory:
 *
 *   CODER warps  (hints -> bits)   A frame is cut into segments of 8 consecutive macroblocks of one row; a thread owns
 *                one segment, a tile is up to NC consecutive segments.
 *     stage    the hints of the tile plus the row above it are copied global -> shared into a swizzled layout
 *              (conflict-free 128-bit reads for threads that each own 64 bytes); the copy of the next tile - of the
 *              next frame when this one ends - is in flight while the current tile is composed.  The frame's
 *              payload bit-offset table arrives by one bulk copy through the TMA unit (cp.async.bulk + mbarrier),
 *              the payload itself is pulled into L2 by a bulk prefetch.
 *     decide   per MB: A/B/C(D) neighbours from shared memory, packed int16 predictor, P_Skip test, mvd, closed-form
 *              Exp-Golomb -> a <= 32-bit code + length in registers; the constant code where a whole warp sits in
 *              uniform motion (replaces src/h264_writer.c:595-641, :369-453, src/bitwriter.c:50-101)
 *     skiprun  SPEC mode only: group prefix-max of "last coded MB" -> mb_skip_run prefixes
 *     scan     group prefix sum of the threads' bit counts -> slice bit offsets
 *     write    a thread whose 8 codes are all short concatenates them branch-free into a 2/4/8-word integer, shifts it
 *              to its slice bit offset and stores whole words into the window; partial first/last words of neighbouring
 *              threads are merged by shuffle
 *     push     everything else - spliced MBs, long codes, separately prefixed codes, the slice trailer - becomes a
 *              16-byte work item in the window's queue
 *     hand-over  when the frame ends (or the window / the queue is full) the window goes to the packer warps through an
 *              mbarrier and the coders go on in the other window
 *
 *   PACKER warps (bits -> Annex-B bytes)
 *     splice   one lane per item: a pre-encoded MB is copied from the payload with funnel shifts (copy_bits,
 *              src/h264_writer.c:228), whole words as plain stores, edge words with shared-memory atomicOr
 *     emit     emulation prevention (src/nal.c:24-50) on 16-byte groups: zero-byte masks, a bit-parallel run-parity
 *              rule, one prefix sum of the insert counts per pass; groups without insertions leave as re-aligned
 *              32/128-bit words, groups with insertions byte-wise by a whole warp (one byte per lane)
 *     clear    the window is zeroed with one st.bulk and handed back
 *
 * The Annex-B start code, the NAL header byte and the host-written slice header are the first bits of the same bit
 * string, so the window image is the output image.
 */
#include <cuda_runtime.h>
#include "mb_codec.cuh"
#include "kernel_util.cuh"
#include "compose_kernels.cuh"

namespace h264scroll {

constexpr int NT = COMPOSE_THREADS;          /* threads per CTA */
constexpr bool FUSED = COMPOSE_FUSED != 0;   /* every thread is coder and packer in turn (no warp specialisation) */
constexpr int NC = FUSED ? NT : COMPOSE_CODER_THREADS;       /* coder threads (warps 0 .. NC/32-1) */
constexpr int NP = FUSED ? NT : NT - COMPOSE_CODER_THREADS;  /* packer threads */
constexpr int NWIN = FUSED ? 1 : 2;          /* window buffers (and queues) per CTA */
constexpr int MPT = 8;                       /* macroblocks per segment / coder thread */
constexpr int NWC = NC / 32, NWP = NP / 32;
constexpr int CH = COMPOSE_WINDOW_WORDS;     /* window: slice words staged per buffer */
constexpr int LCAP = COMPOSE_ITEM_SLOTS;     /* queue items per window hand-over */
constexpr int HBS = COMPOSE_HINT_SLOTS;      /* 64-byte segment slots per staged hint tile */
constexpr int TABCAP = COMPOSE_TAB_WORDS;    /* staged entries of the frame's payload bit-offset table */
constexpr int PAYBUF16 = COMPOSE_PAYBUF16;   /* 16-byte chunks of the frame's payload the packers stage in shared memory */
constexpr u32 FULL = 0xFFFFFFFFu;
constexpr int EG = 4;                        /* emit: 16-byte groups per packer thread per pass */
constexpr int EPASS = EG * NP;               /* emit: groups per pass */
constexpr int BAR_CODER = FUSED ? 0 : 1, BAR_PACKER = FUSED ? 0 : 2;   /* named barriers of the two groups (0 = the whole CTA) */
static_assert(NC % 64 == 0 && NP % 32 == 0 && NP >= 32 && (FUSED || NC + NP == NT), "group sizes");
static_assert(NWC <= 32 && NWP <= 32, "a warp scans the group's warp totals");
static_assert(CH % 4 == 0 && TABCAP % 4 == 0, "16-byte granules");

/* hand-over flags */
#define SLOT_EMIT   1u     /* emit window bytes [0, chunk_bytes) after splicing, then clear the window */
#define SLOT_FIRST  2u     /* the frame's first emit: the emulation-prevention state starts over */
#define SLOT_LAST   4u     /* the frame's last emit: write sizes[f] */
#define SLOT_EXIT   8u     /* no more work */

struct SlotMeta {
    uint8_t *fout;               /* the frame's output slot */
    const u32 *pay_words;        /* the frame's payload */
    u32 *size_ptr;               /* &sizes[f] */
    u32 n_pay_words;
    u32 cap;                     /* output slot capacity */
    u32 buf;                     /* which window */
    u32 n_items;
    u32 flags;
    u32 wbase;                   /* frame word held by window word 0 */
    u32 chunk_bytes;             /* bytes to emit (frame bytes [4 * wbase, 4 * wbase + chunk_bytes)) */
    u32 clear_words;             /* window words the coders may have written */
    u32 carry;                   /* != 0: window words [carry, carry + 4) go on in the other window (OR-ed into its first four) */
    u32 total_bytes;             /* SLOT_LAST: RBSP-side size of the frame's NAL unit (before insertions) */
    int num_refs;
};

struct __align__(128) ComposeSmem {
    u32 win[NWIN][CH + 8];               /* slice words [wbase, wbase + CH) of a frame, MSB first (+ read-ahead pad); zero where unwritten */
    uint4 hbuf[HBS * 4];                 /* the staged hint tile (swizzled 16-byte chunks) */
    uint4 items[NWIN][LCAP];             /* work queues, one per hand-over slot */
    u32 tab[TABCAP];                     /* the frame's payload bit-offset table (bulk copy) */
    uint4 paybuf[PAYBUF16];              /* packers: the frame's pre-encoded payload (bulk copy; what does not fit is read from global memory) */
    SlotMeta meta[2];
    u64 full[2], empty[2];               /* hand-over ring: coders -> packers, packers -> coders */
    u64 tab_full;
    u64 pay_full;
    u32 scan[2][NWC + 1];                /* coders: double buffered, one barrier per group scan */
    u32 escan[2][NWP + 1];               /* packers: the warps' insert totals of an emit pass (double buffered) */
    u32 tile_items[2];                   /* coders: items of the current / next tile */
    u32 zrun_out;                        /* packers: zero bytes closing the bytes emitted so far (rbsp_to_ebsp state) */
};

/* geometry of a frame as the tile loop needs it */
struct FrameGeo {
    const uint2 *fh;      /* the frame's dense hint map */
    int W, H;
    int tpr;              /* segments (threads) per row */
    int n_seg;            /* tpr * H */
    int seg_max;          /* segments per tile */
    u32 tpr_magic;        /* s / tpr == umulhi(s, tpr_magic) for s < 2^25 (tpr > 1) */
};

__device__ __forceinline__ FrameGeo load_geo(const H264GpuFrameDesc *__restrict__ descs, int f,
                                             const H264GpuMbHint *__restrict__ hints) {
    const uint4 *dq = (const uint4 *)(descs + f);
    const uint4 d2 = __ldg(dq + 2), d3 = __ldg(dq + 3);
    FrameGeo g;
    g.fh = (const uint2 *)hints + ((u64)d2.x | ((u64)d2.y << 32));
    g.W = (int)(d3.z & 0xFFFFu);
    g.H = (int)(d3.z >> 16);
    g.tpr = (g.W + MPT - 1) / MPT;
    g.n_seg = g.tpr * g.H;
    g.seg_max = min(NC, HBS - g.tpr - 1);
    g.tpr_magic = 0xFFFFFFFFu / (u32)g.tpr + 1u;
    return g;
}

__device__ __forceinline__ int seg_row(const FrameGeo &g, int s) {
    return g.tpr == 1 ? s : (int)__umulhi((u32)s, g.tpr_magic);
}

/*
 * Stage the hints of segments [s0 - tpr - 1, s0 + n) into buf: slot q holds segment s0 - tpr - 1 + q
 * (8 hints, the tail of a row's last segment stays unwritten), chunks swizzled.  Segments before the
 * frame's first are filled with the "unavailable" hint.  Called by the coder threads (ctid = 0 .. NC-1).
 */
/* frames whose width is not a multiple of 8 macroblocks: chunk by chunk, with its row and column (rare: out of line) */
__device__ __noinline__ void stage_hints_ragged(uint4 *buf, const uint2 *fh, int W, int tpr, u32 tpr_magic, int n_seg, int s_base, int n_chunks) {
    const u32 buf_s = smem_u32(buf);
    for (int c = threadIdx.x; c < n_chunks; c += NC) {
        const u32 pc = hint_swizzle((u32)c);
        const int s = s_base + (c >> 2);
        if (s < 0) { buf[pc] = make_uint4(0u, HINT_Y_UNAVAIL, 0u, HINT_Y_UNAVAIL); continue; }
        if (s >= n_seg) continue;
        const int row = tpr == 1 ? s : (int)__umulhi((u32)s, tpr_magic), x = MPT * (s - row * tpr) + 2 * (c & 3);
        const long long gi = (long long)row * W + x;
        if (x + 1 < W && !(gi & 1)) cp_async16(buf_s + 16u * pc, fh + gi);
        else {
            if (x < W) cp_async8(buf_s + 16u * pc, fh + gi);
            if (x + 1 < W) cp_async8(buf_s + 16u * pc + 8u, fh + gi + 1);
        }
    }
}

/* (out of line: one copy of the copy loop in the instruction stream, whoever calls it) */
__device__ __noinline__ void stage_hints_tile(uint4 *buf, const uint2 *fh, int W, int tpr, u32 tpr_magic, int n_seg, int s0, int n) {
    const int s_base = s0 - tpr - 1;
    const int n_chunks = 4 * (n + tpr + 1);
    if ((W & 7) == 0) {
        /* segments tile the dense map exactly: chunk c of the tile is chunk 4 * s_base + c of the map.  A thread's
           chunks are NC apart, which leaves the swizzle term unchanged: one address each side, then fixed strides */
        const int tid = threadIdx.x;
        const u32 dst = smem_u32(buf) + 16u * hint_swizzle((u32)tid);
        uint4 *dst_g = buf + hint_swizzle((u32)tid);
        const uint4 *src = (const uint4 *)fh + (4 * (long long)s_base + tid);
        const int c_lo = -4 * s_base;                        /* first chunk inside the frame */
        const u32 c_span = (u32)(4 * n_seg);                 /* chunks of the frame */
#pragma unroll
        for (int i = 0; i < (HBS * 4 + NC - 1) / NC; i++) {
            const int c = tid + i * NC;
            if (c < n_chunks) {
                if ((u32)(c - c_lo) < c_span) cp_async16(dst + 16u * (u32)(i * NC), src + i * NC);
                else if (c < c_lo) dst_g[i * NC] = make_uint4(0u, HINT_Y_UNAVAIL, 0u, HINT_Y_UNAVAIL);
            }
        }
    } else stage_hints_ragged(buf, fh, W, tpr, tpr_magic, n_seg, s_base, n_chunks);
}
__device__ __forceinline__ void stage_hints(uint4 *buf, const FrameGeo &g, int s0, int n) {
    stage_hints_tile(buf, g.fh, g.W, g.tpr, g.tpr_magic, g.n_seg, s0, n);
}

/* hint i (0..7) of slot q / 16-byte chunk of a staged tile */
__device__ __forceinline__ uint2 staged_hint(const uint4 *buf, int q, int i) {
    const u32 pc = hint_swizzle((u32)(4 * q + (i >> 1)));
    return ((const uint2 *)buf)[2u * pc + (u32)(i & 1)];
}
__device__ __forceinline__ uint4 staged_chunk(const uint4 *buf, int q, int sub) {
    return buf[hint_swizzle((u32)(4 * q + sub))];
}

/* ===================================================================== packer side */

/* out of line: the backward scan is rare */
__device__ __noinline__ u32 zeros_before_win(const u32 *win, u32 b, u32 zrun_in) { return zeros_before(win, b, zrun_in); }

/* exact insert mask of group q of the window (src/nal.c:24-50 in closed form) */
__device__ __noinline__ u32 group_insert_mask(const u32 *win, int q, uint4 w, int vb, u32 chunk_byte0, u32 zrun_in) {
    const u32 valid = (1u << vb) - 1u;
    u32 z = (zero_nib(w.x) | (zero_nib(w.y) << 4) | (zero_nib(w.z) << 8) | (zero_nib(w.w) << 12)) & valid;
    u32 keep = valid;
    if (chunk_byte0 == 0u && q == 0) { z &= ~0x1Fu; keep &= ~0x1Fu; }   /* start code + NAL header */
    const bool need = ((w.x >> 26) == 0u) && (keep & 1u);                /* a first byte <= 3 may close a zero run of the bytes before */
    if (!(need || (z & (z << 1)))) return 0u;
    u32 rin = 0;
    if (need && (q == 0 || (win[4 * q - 1] & 0xFFu) == 0u)) rin = zeros_before_win(win, 16u * (u32)q, zrun_in);
    const u32 m3 = zero_nib(w.x & 0xFCFCFCFCu) | (zero_nib(w.y & 0xFCFCFCFCu) << 4) |
                   (zero_nib(w.z & 0xFCFCFCFCu) << 8) | (zero_nib(w.w & 0xFCFCFCFCu) << 12);
    return epb_group_mask(z, m3 & keep, vb, rin);
}

/*
 * emit: emulation prevention + store of the window's first chunk_bytes bytes (frame bytes from chunk_byte0 on).
 * Returns the bytes inserted; leaves the closing zero run in sm.zrun_out (read it behind a packer barrier).
 * Called by all packer threads behind a packer barrier.  A pass covers EG 16-byte groups per thread; a warp owns
 * EG * 32 consecutive groups of the pass (round r of warp w = groups (w * EG + r) * 32 + lane: neighbouring lanes hold
 * neighbouring groups, and a round's last group is followed by the next round's first one in the same warp):
 *   masks     the owner computes the exact insert mask of a group that holds a zero byte or starts with a byte <= 3
 *             behind one; every other group inserts nothing
 *   offsets   prefix sums of the counts inside the warp (packed 16-bit fields, EG rounds at once), one exchange of the
 *             warps' totals through shared memory (one barrier per pass)
 *   store     a full group without insertions leaves as aligned or re-aligned 32/128-bit words (the word it shares
 *             with a neighbour is a whole-word store when that neighbour - a bit of the round's ballot - is of the
 *             same kind); groups that insert, the frame's last partial group and everything next to the slot's end
 *             go byte-wise, one byte per lane, two groups per warp and step
 */
__device__ __noinline__ u32 packer_emit(ComposeSmem &sm, const u32 *win, int ptid, u32 chunk_byte0, int chunk_bytes,
                                        u32 ins_total, u32 zrun_in, u32 cap, uint8_t *fout) {
    const int lane = ptid & 31, pw = ptid >> 5;
    u32 *fout32 = (u32 *)fout;
    const u32 ins_in = ins_total;
    const int n_groups = (chunk_bytes + 15) >> 4;
    u32 par = 0;
    for (int q0 = 0; q0 < n_groups; q0 += EPASS, par ^= 1u) {
        /* even with an insertion after every second byte the pass stays clear of the slot's end */
        const bool roomy = chunk_byte0 + 16u * (u32)q0 + 24u * (u32)min(EPASS, n_groups - q0) + ins_total + 4u <= cap;
        const int qw = q0 + pw * (EG * 32) + lane;           /* the thread's group of round 0 */
        u64 insw = 0;                                        /* the groups' insert masks, 16 bits each */
        u32 lo = 0, hi = 0;                                  /* ... and their insert counts (rounds 0, 1 | 2, 3) */
        u32 edge = 0;                                        /* bit 2r: the round's first group is clean, bit 2r + 1: its last one */
#pragma unroll 1
        for (int r = 0; r < EG; r++) {
            const int q = qw + 32 * r;
            const int vb = min(max(chunk_bytes - 16 * q, 0), 16);
            u32 ins = 0u;
            if (vb > 0) {
                const uint4 w = ((const uint4 *)win)[q];
                const u32 hz = ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) |
                               ((w.z - 0x01010101u) & ~w.z) | ((w.w - 0x01010101u) & ~w.w);
                bool sus = (hz & 0x80808080u) != 0u;
                if (!sus && (w.x >> 26) == 0u) sus = q == 0 || (win[4 * q - 1] & 0xFFu) == 0u;
                if (sus) ins = group_insert_mask(win, q, w, vb, chunk_byte0, zrun_in);
                /* the holder of the window's last byte hands the zero run to the next window */
                if (16 * (q + 1) >= chunk_bytes) sm.zrun_out = zeros_before_win(win, (u32)chunk_bytes, zrun_in);
            }
            const u32 cm = __ballot_sync(FULL, vb == 16 && ins == 0u && roomy);
            edge |= ((cm & 1u) | ((cm >> 30) & 2u)) << (2 * r);
            insw |= (u64)ins << (16 * r);
            const u32 c = (u32)__popc(ins) << (16 * (r & 1));
            if (r & 2) hi |= c; else lo |= c;
        }
        /* inclusive prefix sums over the lanes, the warp's totals per round, the warps' totals through shared memory */
        u32 il = lo, ih = hi;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const u32 x = __shfl_up_sync(FULL, il, dd), y = __shfl_up_sync(FULL, ih, dd);
            if (lane >= dd) { il += x; ih += y; }
        }
        const u32 tl = __shfl_sync(FULL, il, 31), th = __shfl_sync(FULL, ih, 31);
        const u32 wsum = (tl & 0xFFFFu) + (tl >> 16) + (th & 0xFFFFu) + (th >> 16);
        if (lane == 0) sm.escan[par][pw] = wsum;
        group_sync<BAR_PACKER, NP>();
        u32 wbase_ins = 0, pass_total = 0;
#pragma unroll
        for (int w = 0; w < NWP; w++) {
            const u32 v = sm.escan[par][w];
            pass_total += v;
            if (w < pw) wbase_ins += v;
        }
        const u64 exw = (u64)(il - lo) | ((u64)(ih - hi) << 32), totw = (u64)tl | ((u64)th << 32);
        u32 round_base = ins_total + wbase_ins;              /* bytes inserted before the round's first group */
#pragma unroll 1
        for (int r = 0; r < EG; r++) {
            const int q = qw + 32 * r;
            const int vb = min(max(chunk_bytes - 16 * q, 0), 16);
            const u32 before = round_base + ((u32)(exw >> (16 * r)) & 0xFFFFu);
            round_base += (u32)(totw >> (16 * r)) & 0xFFFFu;
            const u32 ins = (u32)(insw >> (16 * r)) & 0xFFFFu;
            const u32 D = chunk_byte0 + 16u * (u32)q + before;             /* output byte of the group's first byte */
            const bool clean = vb == 16 && ins == 0u && roomy;
            const u32 cm = __ballot_sync(FULL, clean);
            if (clean) {
                const uint4 w = ((const uint4 *)win)[q];
                const u32 s = (4u - (D & 3u)) & 3u;          /* source bytes before the first aligned word */
                u32 *dst = fout32 + ((D + s) >> 2);
                if (s == 0u) {
                    const uint4 o = make_uint4(__byte_perm(w.x, 0u, 0x0123u), __byte_perm(w.y, 0u, 0x0123u),
                                               __byte_perm(w.z, 0u, 0x0123u), __byte_perm(w.w, 0u, 0x0123u));
                    if ((D & 15u) == 0u) *(uint4 *)dst = o;
                    else { dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w; }
                } else {
                    /* re-aligned words; the bytes shared with a neighbour are whole-word stores only when the neighbour is
                       clean too: the lanes beside this one, the previous round's last / the next round's first group at the
                       warp's ends (nobody beyond the warp's span: those few bytes go one by one) */
                    const bool prev_clean = lane > 0 ? (cm >> (lane - 1)) & 1u : (r > 0 && ((edge >> (2 * r - 1)) & 1u));
                    const bool next_clean = lane < 31 ? (cm >> (lane + 1)) & 1u : (r < EG - 1 && ((edge >> (2 * r + 2)) & 1u));
                    const u32 wn = win[4 * q + 4];
                    const u32 sel = realign_selector(s);
                    dst[0] = __byte_perm(w.x, w.y, sel);
                    dst[1] = __byte_perm(w.y, w.z, sel);
                    dst[2] = __byte_perm(w.z, w.w, sel);
                    const u32 o3 = __byte_perm(w.w, wn, sel);
                    if (next_clean) dst[3] = o3;
                    else {                                   /* (1..3 bytes, no loop) */
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
            /* the others: the warp takes them two at a time, a byte per lane */
            u32 m = __ballot_sync(FULL, vb > 0 && !clean);
            while (m) {
                const int l0 = __ffs((int)m) - 1;
                m &= m - 1u;
                int l1 = -1;
                if (m) { l1 = __ffs((int)m) - 1; m &= m - 1u; }
                const int l = (lane < 16 || l1 < 0) ? l0 : l1;
                const u32 Dl = __shfl_sync(FULL, D, l), il2 = __shfl_sync(FULL, ins, l);
                const int vl = __shfl_sync(FULL, vb, l), ql = __shfl_sync(FULL, q, l);
                const int j = lane & 15;
                if (j < vl && (lane < 16 || l1 >= 0)) {
                    const u32 byte = (win[4 * ql + (j >> 2)] >> (24 - 8 * (j & 3))) & 0xFFu;
                    const u32 pos = Dl + (u32)j + (u32)__popc(il2 & ((2u << j) - 1u));
                    if ((il2 >> j) & 1u) { if (pos - 1u < cap) fout[pos - 1u] = 3; }
                    if (pos < cap) fout[pos] = (uint8_t)byte;
                }
            }
        }
        ins_total += pass_total;
    }
    return ins_total - ins_in;
}

/* splice: one lane per queue item; a spliced macroblock whose payload words (and the read-ahead word) are all staged
   reads shared memory */
__device__ __forceinline__ void packer_splice(const uint4 *items, u32 n_items, int ptid, u32 *win, u32 wbase, int num_refs,
                                           const u32 *pw, u32 n_words, const u32 *ps, u32 staged) {
    for (u32 i = (u32)ptid; i < n_items; i += NP) {
        const uint4 q = items[i];
        Item it; it.x = q.x; it.y = q.y; it.z = q.z; it.w = q.w;
        const bool in_smem = (q.x >> 30) == IT_DYN && ((q.z + q.w + 31u) >> 5) + 1u < staged;
        consume_item(it, win, wbase, (u32)CH, num_refs, in_smem ? ps : pw, n_words);
    }
}

/* what the packing side carries from window to window */
struct PackState {
    u32 ins_total;                     /* emulation-prevention bytes inserted so far in the current frame */
    u32 zrun;                          /* zero bytes closing what the frame has emitted so far */
    const u32 *frame_id;               /* the frame whose payload is staged (its sizes[] entry names it) */
    const u32 *pay_smem;               /* payload word i of that frame = pay_smem[i] for i < pay_staged */
    u32 pay_staged, pay_par;
};
__device__ __forceinline__ void pack_state_init(PackState &st) {
    st.ins_total = 0; st.zrun = 0; st.frame_id = nullptr; st.pay_smem = nullptr; st.pay_staged = 0; st.pay_par = 0;
}

/*
 * A frame's payload comes to shared memory with one bulk copy of the whole 16-byte chunks that hold it (what does not
 * fit, and the odd bytes at its ends, are read from global memory).  `issuer` issues the copy; everybody learns what is
 * staged.  Nobody may still read the previous frame's payload (a frame's last hand-over is an emit, with barriers
 * behind its splice).
 */
__device__ __forceinline__ void stage_payload(ComposeSmem &sm, PackState &st, const u32 *frame_id, const u32 *pay_words, u32 n_pay_words, bool issuer) {
    st.frame_id = frame_id;
    st.pay_staged = 0;
    st.pay_par &= 1u;                                        /* (bit 1: this frame's copy has been waited for) */
    if (PAYBUF16 > 0 && pay_words != nullptr) {
        const uintptr_t a0 = (uintptr_t)pay_words & ~(uintptr_t)15;
        const u32 w_skip = (u32)(((uintptr_t)pay_words - a0) >> 2);
        const u32 n16 = min((w_skip + n_pay_words) >> 2, (u32)PAYBUF16);
        if (4u * n16 > w_skip) {
            if (issuer) {
                mbar_arrive_expect_tx(&sm.pay_full, 16u * n16);
                bulk_g2s(sm.paybuf, (const void *)a0, 16u * n16, &sm.pay_full);
            }
            st.pay_staged = 4u * n16 - w_skip;
            st.pay_smem = (const u32 *)sm.paybuf + w_skip;
        }
    }
}
/* wait for the staged payload (once per frame: pay_par flips) */
__device__ __forceinline__ void await_payload(ComposeSmem &sm, PackState &st) {
    if (st.pay_staged && !(st.pay_par & 2u)) { mbar_wait(&sm.pay_full, st.pay_par & 1u); st.pay_par = (st.pay_par ^ 1u) | 2u; }
}

/* one hand-over: splice the queue into the window, then (SLOT_EMIT) emit and clear it.  Called by all NP packing
   threads; the barriers are the packing group's */
__device__ __noinline__ void process_slot(ComposeSmem &sm, const SlotMeta &m, const uint4 *items, int ptid, PackState &st) {
    const u32 flags = m.flags;
    u32 *win = sm.win[m.buf];
    const u32 wbase = m.wbase;
    if (!FUSED && m.size_ptr != st.frame_id)                 /* (the fused coder stages at frame start) */
        stage_payload(sm, st, m.size_ptr, m.pay_words, m.n_pay_words, ptid == 0);
    await_payload(sm, st);
    packer_splice(items, m.n_items, ptid, win, wbase, m.num_refs, m.pay_words, m.n_pay_words, st.pay_smem, st.pay_staged);
    if (flags & SLOT_EMIT) {
        group_sync<BAR_PACKER, NP>();                        /* every spliced bit is in */
        u32 cval = 0;                                        /* the frame goes on behind this window: its last words move on */
        if (m.carry && ptid < 4) {
            cval = win[m.carry + (u32)ptid];
            if (!FUSED && cval) atomicOr(&sm.win[(m.buf ^ 1u) & (NWIN - 1)][ptid], cval);
        }
        if (flags & SLOT_FIRST) { st.ins_total = 0; st.zrun = 0; }
        const u32 chunk_bytes = m.chunk_bytes;
        st.ins_total += packer_emit(sm, win, ptid, wbase << 2, (int)chunk_bytes, st.ins_total, st.zrun, m.cap, m.fout);
        group_sync<BAR_PACKER, NP>();                        /* window fully read */
        if (chunk_bytes > 0u) st.zrun = sm.zrun_out;
        const u32 cw = (m.clear_words + 3u) & ~3u;
#if COMPOSE_BULK_CLEAR
        if (ptid == 0 && cw) smem_zero_bulk(win, 4u * cw);
#else
        for (u32 i = (u32)ptid; i < cw / 4u; i += NP) ((uint4 *)win)[i] = make_uint4(0u, 0u, 0u, 0u);
#endif
        if ((flags & SLOT_LAST) && ptid == 0) *m.size_ptr = m.total_bytes + st.ins_total;
        if (FUSED && m.carry) {                              /* (one window: the words come back in at its start) */
            group_sync<BAR_PACKER, NP>();
            if (ptid < 4 && cval) atomicOr(&win[ptid], cval);
        }
    }
}

/* the packer group: takes windows off the ring until it is told to leave */
__device__ __forceinline__ void packer_main(ComposeSmem &sm, int ptid) {
    PackState st;
    pack_state_init(st);
    for (int s = 0;; s++) {
        mbar_wait(&sm.full[s & 1], (u32)(s >> 1) & 1u);
        const SlotMeta &m = sm.meta[s & 1];
        if (m.flags & SLOT_EXIT) break;
        process_slot(sm, m, sm.items[s & (NWIN - 1)], ptid, st);
        mbar_arrive(&sm.empty[s & 1]);
    }
}

/* ===================================================================== coder side */

/*
 * The concatenation path of a thread whose macroblocks are all MBK_SHORT / MBK_NONE: bits [O, O + Ls)
 * of the slice.  Stores the whole words it owns; hands back the partial head word H (word O >> 5, when
 * O is not word aligned) and the last word T (word (O + Ls - 1) >> 5) unless that is the head word.
 */
template <int NWc>
__device__ __forceinline__ void simple_words(const u32 (&code)[MPT], const u32 (&len)[MPT], u32 O, u32 Ls,
                                             u32 *win, u32 wbase, u32 &H, u32 &T, bool &has_t) {
    u32 a[NWc];
#pragma unroll
    for (int i = 0; i < NWc; i++) a[i] = 0u;
#pragma unroll
    for (int k = 0; k < MPT; k++) concat_append<NWc>(a, code[k], len[k], k);
    const u32 E = O + Ls;
    const u32 j0 = O >> 5, s = O & 31u, jl = (E - 1u) >> 5;
    u32 b[NWc + 1];
    concat_place<NWc>(a, (32u - (E & 31u)) & 31u, b);
    const u32 lo_idx = j0 + (s ? 1u : 0u);                   /* first word this thread stores itself */
    H = 0u;
#pragma unroll
    for (int i = 0; i < NWc; i++) {
        const u32 idx = jl - (u32)(NWc - i);
        if ((int)(idx - lo_idx) >= 0 && idx - wbase < (u32)CH) win[idx - wbase] = b[i];
        if (idx == j0) H = b[i];
    }
    has_t = !(jl == j0 && s);
    T = b[NWc];
    if (!has_t) H = b[NWc];
    if (!s) H = 0u;
}

/*
 * The 8 macroblocks of a segment -> codes, lengths, kinds (4 bits each).  GEN = 0 is the reference mode
 * (skip OFF, reference predictor) with everything mode dependent folded away; GEN = 1 takes the flags.
 * A macroblock in uniform motion - every available neighbour carries its own hint - takes the constant
 * zero-mvd code; the vote keeps the warp on that path as long as no lane needs the predictor
 * (spliced macroblocks do not: their length comes from the offset table).
 */
template <int GEN>
__device__ __forceinline__ u32 decide_segment(const uint2 (&own)[MPT + 1], const uint2 (&abv)[MPT + 2], int nmb, bool row0,
                                              int num_refs, const u32 *tab, u32 (&code)[MPT], u32 (&len)[MPT],
                                              int skip_spec_in = 0, int pred_spec_in = 0) {
    const int skip_spec = GEN ? skip_spec_in : 0, pred_spec = GEN ? pred_spec_in : 0;
    u32 kinds = 0;
#pragma unroll
    for (int k = 0; k < MPT; k++) {
        const uint2 h = own[1 + k], a = own[k], b = abv[k + 1], c = abv[k + 2];
        const bool present = k < nmb;
        const bool dyn = (h.y & 0xFF00u) != 0u;
        bool uni;
        if (skip_spec) {
            /* P_Skip looks at A and B themselves: all three neighbours must carry the MB's own hint */
            uni = ((h.x ^ a.x) | (h.y ^ a.y) | (h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
        } else if (row0) {
            /* top row: A alone decides (src/h264_writer.c:407-417); no neighbour at all predicts 0 */
            uni = ((h.x ^ a.x) | (h.y ^ a.y)) == 0u;
        } else {
            /* B == C == the MB's own hint: two refs match, and both medians of (a, h, h) are h whatever A is */
            uni = ((h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
        }
        u32 kd = MBK_NONE;
        code[k] = 0u; len[k] = 0u;
        if (__all_sync(FULL, uni || dyn || !present)) {
            if (present && !dyn) kd = decide_mb_uniform(h.y, num_refs, skip_spec, code[k], len[k]);
        } else if (present && !dyn) {
            kd = decide_mb(h.x, h.y, a.x, a.y, b.x, b.y, c.x, c.y, num_refs, pred_spec, skip_spec, code[k], len[k]);
        }
        if (present && dyn) {
            const u32 di = h.y >> 16;
            const u32 b0 = tab[di];                          /* (generic loads: the table is staged in shared memory when it fits) */
            len[k] = tab[di + 1] - b0 + (skip_spec ? 0u : 1u);
            code[k] = b0;
            kd = MBK_DYN;
        }
        kinds |= kd << (4 * k);
    }
    return kinds;
}

/* frames whose flags ask for SPEC skip / SPEC prediction belong to the SPEC = true instantiation, all others to the
   SPEC = false one: the reference-mode kernel carries none of the SPEC code (the kernel is sensitive to the size of
   its instruction stream).  Both are launched over the whole batch; each walks only its own frames. */
__device__ __forceinline__ bool frame_is_spec(const H264GpuFrameDesc *__restrict__ descs, int f) {
    return ((__ldg((const u32 *)(descs + f) + 15) >> 16) & (H264GPU_SKIP_SPEC | H264GPU_PRED_SPEC)) != 0u;
}
/* first frame f, f + step, f + 2 step ... that belongs to this instantiation; called by whole warps with the same
   arguments: the lanes look at 32 candidates at a time */
template <bool SPEC>
__device__ __forceinline__ int next_own_frame(const H264GpuFrameDesc *__restrict__ descs, int f, int n_frames, int step) {
    const int lane = threadIdx.x & 31;
    while (f < n_frames) {
        const long long mine = (long long)f + (long long)lane * step;
        const bool own = mine < n_frames && frame_is_spec(descs, (int)mine) == SPEC;
        const u32 m = __ballot_sync(FULL, own);
        if (m) return f + (__ffs((int)m) - 1) * step;
        if ((long long)f + 32ll * step >= n_frames) return n_frames;
        f += 32 * step;
    }
    return f;
}

/* the coders' view of the hand-over ring (every coder thread tracks it identically) */
struct Ring {
    int slot;             /* the slot being filled */
    int confirmed;        /* every slot up to here is known to be consumed */
};
__device__ __forceinline__ void ring_wait(ComposeSmem &sm, Ring &r, int s) {
    if (s > r.confirmed) { mbar_wait(&sm.empty[s & 1], (u32)(s >> 1) & 1u); r.confirmed = s; }
}

template <bool SPEC>
__device__ __forceinline__ void coder_main(ComposeSmem &sm, int f, const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                                           const H264GpuMbHint *__restrict__ hints,
                                           const H264GpuDynDesc *__restrict__ dyns,
                                           const u32 *__restrict__ dyn_tab,
                                           const uint8_t *__restrict__ payload,
                                           uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    const int tid = threadIdx.x, lane = tid & 31, gw = tid >> 5;
    u32 spar = 0;                      /* which sm.scan buffer the next group scan uses */
    u32 ipar = 0;                      /* which sm.tile_items counter the current tile uses */
    Ring ring; ring.slot = 0; ring.confirmed = -1;   /* (warp-specialised build only) */
    PackState pst;                     /* (fused build only: this thread is a packer too) */
    pack_state_init(pst);
    u32 buf = 0;                       /* the window being filled */
    u32 items_used = 0;                /* items queued in the slot being filled */
    u32 tab_par = 0;                   /* phase of sm.tab_full the next staged table completes */

    FrameGeo geo = load_geo(descs, f, hints);
    stage_hints(sm.hbuf, geo, 0, min(geo.seg_max, geo.n_seg));
    cp_async_commit();

    while (f < n_frames) {
        /* ---- frame setup: the descriptor is read by every thread (one broadcast line) ------ */
        const uint4 *dq = (const uint4 *)(descs + f);
        const uint4 d2 = __ldg(dq + 2), d3 = __ldg(dq + 3);
        const u32 dyn_id = d3.y;
        const int W = geo.W;
        const u32 hdr_total = 40u + (d3.w & 0xFFu);
        const int num_refs = (int)((d3.w >> 8) & 0xFFu);
        const int skip_spec = SPEC ? ((d3.w >> 16) & H264GPU_SKIP_SPEC) != 0 : 0;
        const int pred_spec = SPEC ? ((d3.w >> 16) & H264GPU_PRED_SPEC) != 0 : 0;
        const u32 nal_header = d3.w >> 24;
        const u32 *tab = nullptr;          /* the frame's payload bit-offset table as decide reads it */
        bool tab_staged = false;
        if (dyn_id != H264GPU_NO_DYN) {
            const uint4 *yq = (const uint4 *)(dyns + dyn_id);
            const uint4 y0 = __ldg(yq), y1 = __ldg(yq + 1);
            tab = dyn_tab + ((u64)y0.z | ((u64)y0.w << 32));
            /* the table (n_mbs + 1 words) goes to shared memory with one bulk copy of the 16-byte chunks that hold it;
               every coder is past its last read of the previous frame's table (they all arrived at that frame's last
               scan barrier, which this thread has left) */
            const uintptr_t a0 = (uintptr_t)tab & ~(uintptr_t)15;
            const u32 t_skip = (u32)(((uintptr_t)tab - a0) >> 2);
            const u32 t_bytes = (4u * (t_skip + y1.x + 1u) + 15u) & ~15u;
            tab_staged = y1.x < (u32)(TABCAP - 4);
            if (tab_staged) {
                if (tid == 0) {
                    mbar_arrive_expect_tx(&sm.tab_full, t_bytes);
                    bulk_g2s(sm.tab, (const void *)a0, t_bytes, &sm.tab_full);
                }
                tab = sm.tab + t_skip;
            }
            /* the payload is read by the packers straight from L2: send it on its way there now */
            const uint8_t *pa = payload + ((u64)y0.x | ((u64)y0.y << 32));
            if (FUSED) stage_payload(sm, pst, sizes + f, (const u32 *)pa, (y1.y + 3u) >> 2, tid == 0);
            else if (tid == 32 && y1.y) {
                const uintptr_t p0 = (uintptr_t)pa & ~(uintptr_t)15;
                const u64 p_bytes = (((uintptr_t)pa - p0) + (u64)y1.y + 15ull) & ~15ull;
                bulk_prefetch_l2((const void *)p0, (u32)(p_bytes < (1ull << 20) ? p_bytes : (1ull << 20)));
            }
        } else if (FUSED) stage_payload(sm, pst, sizes + f, nullptr, 0u, false);
        const int kc = (W - 1) & 7;                            /* position of a row's last MB inside its segment */

        /* frame state every thread tracks identically (all of it follows from group-wide totals) */
        u32 bitpos = hdr_total;        /* slice bits composed so far (incl. the 40-bit Annex-B prefix) */
        int last_coded = -1;           /* raster index of the last non-skipped MB */
        u32 wbase = 0;                 /* frame word held by window word 0; words below it have left for the packers */
        bool first_emit = true;

        const int f_next = next_own_frame<SPEC>(descs, f + (int)gridDim.x, n_frames, (int)gridDim.x);
        FrameGeo geo_next = geo;
        if (f_next < n_frames) geo_next = load_geo(descs, f_next, hints);

        /* a window (+ its queue) goes to the packers; thread 0 reads the frame's constants again (they are not worth
           registers in every thread).  Slots alternate between the windows except behind a splice-only hand-over, whose
           completion the coders wait for: whoever fills slot s may touch its window and its queue once slot s - 2 is done */
        auto hand_over = [&](u32 flags, u32 chunk_bytes, u32 clear_words, u32 total_bytes, u32 carry = 0u) {
            if (tid == 0) {
                SlotMeta &m = sm.meta[FUSED ? 0 : ring.slot & 1];
                const uint4 e2 = __ldg(dq + 2), e3 = __ldg(dq + 3);
                m.fout = out + ((u64)e2.z | ((u64)e2.w << 32)); m.cap = e3.x;
                m.pay_words = nullptr; m.n_pay_words = 0;
                if (e3.y != H264GPU_NO_DYN) {
                    const uint4 *yq = (const uint4 *)(dyns + e3.y);
                    const uint4 y0 = __ldg(yq), y1 = __ldg(yq + 1);
                    m.pay_words = (const u32 *)(payload + ((u64)y0.x | ((u64)y0.y << 32)));
                    m.n_pay_words = (y1.y + 3u) >> 2;
                }
                m.size_ptr = sizes + f;
                m.buf = buf; m.n_items = items_used; m.flags = flags;
                m.wbase = wbase; m.chunk_bytes = chunk_bytes; m.clear_words = clear_words; m.total_bytes = total_bytes;
                m.carry = carry; m.num_refs = num_refs;
            }
            if (FUSED) {                                            /* every thread turns packer for this window */
                __syncthreads();
                process_slot(sm, sm.meta[0], sm.items[0], tid, pst);
                __syncthreads();
            } else {
                mbar_arrive(&sm.full[ring.slot & 1]);
                ring.slot++;
            }
            items_used = 0;
        };
        /* the window to fill and the queue of the slot to fill are free again */
        auto acquire = [&]() { if (!FUSED) ring_wait(sm, ring, ring.slot - 2); };

        for (int s0 = 0; s0 < geo.n_seg;) {
            const int n = min(geo.seg_max, geo.n_seg - s0);
            const bool last = (s0 + n >= geo.n_seg);
            const bool tile0 = (s0 == 0);
            const uint4 *hb = sm.hbuf;

            /* ---- stage: this tile's hints have landed ---------- */
            cp_async_wait_all();
            group_sync<BAR_CODER, NC>();
            if (tile0 && tab_staged) mbar_wait(&sm.tab_full, tab_par);

            /* ---- load: own segment, the one above, the hints beside them ------------------------ */
            const int s = s0 + tid;
            const bool active = tid < n;
            const int row = seg_row(geo, active ? s : s0), col = (active ? s : s0) - row * geo.tpr;
            const int nmb = active ? min(MPT, W - MPT * col) : 0;
            const int g0 = row * W + MPT * col;                 /* raster index of the segment's first MB */
            const bool first = col == 0, lastseg = col == geo.tpr - 1;
            const int qo = tid + geo.tpr + 1, qa = tid + 1;     /* slots of the own segment and of the one above */

            uint2 own[MPT + 1];                                 /* own[0] = left neighbour of the first MB */
            uint2 abv[MPT + 2];                                 /* abv[j] = hint above-left of MB j */
            if (active) {
#pragma unroll
                for (int k = 0; k < MPT / 2; k++) {
                    const uint4 v = staged_chunk(hb, qo, k), u = staged_chunk(hb, qa, k);
                    own[1 + 2 * k] = make_uint2(v.x, v.y); own[2 + 2 * k] = make_uint2(v.z, v.w);
                    abv[1 + 2 * k] = make_uint2(u.x, u.y); abv[2 + 2 * k] = make_uint2(u.z, u.w);
                }
                own[0] = staged_hint(hb, qo - 1, 7);
                abv[0] = staged_hint(hb, qa - 1, 7);
                abv[MPT + 1] = staged_hint(hb, qa + 1, 0);
                if (first) { own[0] = make_uint2(0u, HINT_Y_UNAVAIL); abv[0] = make_uint2(0u, HINT_Y_UNAVAIL); }
                if (lastseg) {                                  /* C of the row's last MB is D (src/h264_writer.c:391-399) */
                    if (kc == 7) abv[9] = abv[7];
                    else {
#pragma unroll
                        for (int j = 0; j < MPT - 1; j++)
                            if (kc == j) abv[j + 2] = abv[j];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k <= MPT; k++) own[k] = make_uint2(0u, HINT_Y_UNAVAIL);
#pragma unroll
                for (int j = 0; j < MPT + 2; j++) abv[j] = make_uint2(0u, HINT_Y_UNAVAIL);
            }

            /* ---- decide + code ------------------------------------------------ */
            u32 code[MPT], len[MPT];
            u32 kinds;
            if (!SPEC) kinds = decide_segment<0>(own, abv, nmb, row == 0, num_refs, tab, code, len);
            else kinds = decide_segment<1>(own, abv, nmb, row == 0, num_refs, tab, code, len, skip_spec, pred_spec);

            /* ---- mb_skip_run prefixes (SPEC skip; with skip OFF the constant bit is already in) ---- */
            int last_coded_total = last_coded;
            int prev0 = 0;
            if (skip_spec) {
                int lc = -1;
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if ((kinds >> (4 * k)) & 0xFu) lc = g0 + k;
                int blk_max;
                prev0 = max(group_excl_scan_max<NWC, BAR_CODER>(lc, (int *)sm.scan[spar], gw, blk_max), last_coded);
                spar ^= 1u;
                last_coded_total = max(blk_max, last_coded);
                attach_runs(code, len, kinds, g0, prev0);
            }

            /* ---- slice trailer: trailing mb_skip_run + rbsp_stop_one_bit ---------- */
            const bool owns_trailer = last && tid == n - 1;
            u32 tr_code = 1u, tr_len = 1u;
            if (owns_trailer && skip_spec) {
                const u32 run_end = (u32)(W * geo.H - 1 - last_coded_total);
                if (run_end > 0u) {
                    tr_code = ((run_end + 1u) << 1) | 1u;
                    tr_len = 2u * flo32(run_end + 1u) + 2u;
                }
            }

            /* ---- bit count of the thread, queue slots ---------------------------- */
            u32 Ls = 0;
#pragma unroll
            for (int k = 0; k < MPT; k++) Ls += len[k];
            const bool cplx = kinds_complex(kinds);
            u32 n_it = 0;
            if (cplx) {
#pragma unroll
                for (int k = 0; k < MPT; k++) n_it += len[k] ? 1u : 0u;
            }
            if (owns_trailer) n_it++;
            u32 ibase = 0;
            if (n_it) ibase = atomicAdd(&sm.tile_items[ipar], n_it);
            const u32 L = Ls + (owns_trailer ? tr_len : 0u);

            /* ---- bit offsets: prefix sum of the threads' bit counts ----------------- */
            u32 total;
            const u32 O = bitpos + group_excl_scan<NWC, BAR_CODER>(L, sm.scan[spar], gw, total);
            spar ^= 1u;
            const u32 total_items = sm.tile_items[ipar];
            /* ---- stage: everybody is past its loads from the hint tile: start the copy of the next one ---- */
            if (!last || f_next < n_frames) {
                const FrameGeo &gn = last ? geo_next : geo;
                const int sn = last ? 0 : s0 + n;
                stage_hints(sm.hbuf, gn, sn, min(gn.seg_max, gn.n_seg - sn));
            }
            cp_async_commit();
            if (tid == 0) sm.tile_items[ipar ^ 1u] = 0;            /* the next tile's counter; nobody adds to it before the next barrier */
            const u32 E = O + L;
            const u32 end = bitpos + total;
            const u32 total_bytes = (end + 7u) >> 3;                       /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* frame words [.., w_stop) are complete after this tile */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end; /* last word this tile writes to */
            const u32 warp_end = __shfl_sync(FULL, E, 31);

            /* ---- the tile does not fit behind what the window (or its queue) already holds: send that off first ---- */
            if (!tile0 && (w_last - wbase >= (u32)CH || items_used + total_items > (u32)LCAP)) {
                const u32 wf_al = (bitpos >> 5) & ~3u;             /* 16-byte group of the tile's first word */
                if (wf_al > wbase) {
                    /* the window's whole groups leave; its last (up to four) words go on in the other window, which must be
                       free by now.  The packers move them once the queued items are spliced (the last macroblocks of the tile
                       before may be among them) - OR-ed in, as the tile's own writes to those words are */
                    const u32 src = wf_al - wbase;
                    const u32 nb = buf ^ 1u;
                    if (!FUSED) ring_wait(sm, ring, ring.slot - 1);
                    hand_over(SLOT_EMIT | (first_emit ? SLOT_FIRST : 0u), src << 2, src + 4u, 0u, src);
                    first_emit = false;
                    if (!FUSED) buf = nb;
                    wbase = wf_al;
                } else if (items_used) {
                    /* nothing whole to emit yet, but the queue must be drained: splice only, and wait for it */
                    hand_over(0u, 0u, 0u, 0u);
                    if (!FUSED) ring_wait(sm, ring, ring.slot - 1);
                }
            }

            for (;;) {
                acquire();
                u32 *win = sm.win[buf];
                /* ---- header: start code, NAL header byte, slice header bits ---------- */
                if (tile0 && tid < 10 && 32u * (u32)tid < hdr_total) {
                    const uint4 h0 = __ldg(dq), h1 = __ldg(dq + 1);
                    const u64 hdr[4] = {(u64)h0.x | ((u64)h0.y << 32), (u64)h0.z | ((u64)h0.w << 32),
                                        (u64)h1.x | ((u64)h1.y << 32), (u64)h1.z | ((u64)h1.w << 32)};
                    u32 v = header_bits32(hdr, nal_header, 32u * (u32)tid);
                    const u32 left = hdr_total - 32u * (u32)tid;
                    if (left < 32u) v &= ~(FULL >> left);
                    if ((u32)tid - wbase < (u32)CH && v) atomicOr(&win[(u32)tid - wbase], v);
                }

                /* ---- write: concatenation path + merge of the partial words between threads ---------- */
                {
                    const u32 Lc = cplx ? 0u : Ls;                         /* bits through the concatenation path */
                    const u32 Ec = O + Lc;
                    u32 H = 0u, T = 0u;
                    bool has_t = false;
                    const u32 lmax = __reduce_max_sync(FULL, Lc);
                    if (lmax <= 64u) { if (Lc) simple_words<2>(code, len, O, Lc, win, wbase, H, T, has_t); }
                    else if (lmax <= 128u) { if (Lc) simple_words<4>(code, len, O, Lc, win, wbase, H, T, has_t); }
                    else { if (Lc) simple_words<8>(code, len, O, Lc, win, wbase, H, T, has_t); }
                    const u32 key = O >> 5;
                    u32 tj = (Ec - 1u) >> 5;
                    if (!Lc) {
                        /* nothing through this path: the thread still closes the word its range ends in for the heads behind it */
                        tj = E >> 5;
                        has_t = L != 0u && (E & 31u) != 0u && (tj > key || (O & 31u) == 0u);
                    }
                    const bool plain = Lc != 0u && (has_t || (O & 31u) == 0u) && ((O & 31u) == 0u || tj > key);
                    if (__all_sync(FULL, plain)) {
                        /* every string reaches past its head word: a word is shared by two neighbours at most */
                        u32 v = T | __shfl_down_sync(FULL, H, 1);
                        if (lane == 31) v = T;
                        if (tj - wbase < (u32)CH) {
                            if (lane == 31) atomicOr(&win[tj - wbase], v);   /* the next warp's (or tile's) first head lands here */
                            else win[tj - wbase] = v;
                        }
                        if (lane == 0 && H != 0u && key - wbase < (u32)CH) atomicOr(&win[key - wbase], H);
                    } else {
                        /* general case: segmented suffix OR of the heads that share a word */
                        u32 X = H;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const u32 v = __shfl_down_sync(FULL, X, d);
                            const u32 kk = __shfl_down_sync(FULL, key, d);
                            if (lane + d < 32 && kk == key) X |= v;
                        }
                        const u32 Xn = __shfl_down_sync(FULL, X, 1);
                        const u32 kn = __shfl_down_sync(FULL, key, 1);
                        if (has_t && tj - wbase < (u32)CH) {
                            u32 v = T;
                            if (lane < 31 && kn == tj) v |= Xn;
                            if ((warp_end >> 5) == tj) atomicOr(&win[tj - wbase], v);   /* later warps (or the next tile) add to this word */
                            else win[tj - wbase] = v;
                        }
                        if (lane == 0 && X != 0u && key - wbase < (u32)CH) atomicOr(&win[key - wbase], X);
                    }
                }

                /* ---- push: this tile's items into the slot's queue; a tile with more items than the queue has room for
                        goes in rounds, each sent ahead as a splice-only hand-over ----------- */
                for (u32 r_lo = 0;;) {
                    const u32 r_hi = min(total_items, r_lo + ((u32)LCAP - items_used));
                    if (n_it && ibase < r_hi && ibase + n_it > r_lo) {
                        uint4 *queue = sm.items[FUSED ? 0 : ring.slot & 1] + items_used;
                        u32 idx = ibase, bit = O;
                        int prev = prev0;
                        if (cplx) {
#pragma unroll
                            for (int k = 0; k < MPT; k++) {
                                const u32 kd = (kinds >> (4 * k)) & 0xFu;
                                if (len[k]) {
                                    const u32 pre = skip_spec ? (u32)(g0 + k - prev) : 1u;
                                    prev = g0 + k;
                                    if (idx - r_lo < r_hi - r_lo) {
                                        uint2 h = make_uint2(0u, 0u);
                                        if (kd == MBK_LONG) h = __ldg(geo.fh + g0 + k);   /* the staged tile is being replaced */
                                        const Item it = make_item(kd, code[k], len[k], bit, pre, h.x, h.y);
                                        queue[idx - r_lo] = make_uint4(it.x, it.y, it.z, it.w);
                                    }
                                    idx++;
                                    bit += len[k];
                                }
                            }
                        } else bit += Ls;
                        if (owns_trailer && idx - r_lo < r_hi - r_lo)
                            queue[idx - r_lo] = make_uint4(bit | (IT_BITS << 30), 0u, tr_code, tr_len);
                    }
                    items_used += r_hi - r_lo;
                    if (r_hi >= total_items) break;
                    hand_over(0u, 0u, 0u, 0u);                             /* (the same window goes on) */
                    if (!FUSED) ring_wait(sm, ring, ring.slot - 1);
                    r_lo = r_hi;
                }

                const bool fits = w_last - wbase < (u32)CH;                /* the tile's last word lies inside the window */
                if (!(last || !fits)) break;                               /* keep filling the window with the next tile */

                /* ---- the window is full or the frame ends: frame words [wbase, e_hi) go to the packers --------- */
                const u32 e_hi = fits ? w_stop : wbase + CH;
                const u32 chunk_byte0 = wbase << 2;
                const u32 chunk_bytes = e_hi > wbase ? ((fits ? min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0u;
                hand_over(SLOT_EMIT | (first_emit ? SLOT_FIRST : 0u) | (fits ? SLOT_LAST : 0u), chunk_bytes,
                          min((u32)CH, w_last - wbase + 1u), total_bytes);
                first_emit = false;
                if (!FUSED) buf ^= 1u;
                if (fits) break;                                           /* (frame done) */
                wbase += CH;                                               /* a full window has left; the tile goes on in the next one */
            }

            bitpos = end;
            last_coded = last_coded_total;
            ipar ^= 1u;
            s0 += n;
        }
        if (tab_staged) tab_par ^= 1u;
        f = f_next;
        geo = geo_next;
    }
    cp_async_wait_all();
    if (!FUSED) {                                                          /* tell the packers to leave */
        ring_wait(sm, ring, ring.slot - 2);
        if (tid == 0) sm.meta[ring.slot & 1].flags = SLOT_EXIT;
        mbar_arrive(&sm.full[ring.slot & 1]);
    }
}

#include "coder_lean.cuh"

template <bool SPEC>
__global__ void __launch_bounds__(NT, COMPOSE_MIN_CTAS)
compose_frames_kernel(const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                      const H264GpuMbHint *__restrict__ hints,
                      const H264GpuDynDesc *__restrict__ dyns,
                      const u32 *__restrict__ dyn_tab,
                      const uint8_t *__restrict__ payload,
                      uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ComposeSmem &sm = *reinterpret_cast<ComposeSmem *>(smem_raw);
    const int tid = threadIdx.x;

    const int f = next_own_frame<SPEC>(descs, blockIdx.x, n_frames, (int)gridDim.x);
    if (f >= n_frames) return;                                 /* (the whole CTA) */

    {   /* the windows are zero wherever nothing was written: cleared once here, then the packers clear what was used */
        uint4 *w4 = (uint4 *)sm.win;
        for (int i = tid; i < NWIN * (CH + 8) / 4; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) {
            mbar_init(&sm.full[0], NC); mbar_init(&sm.full[1], NC);
            mbar_init(&sm.empty[0], NP); mbar_init(&sm.empty[1], NP);
            mbar_init(&sm.tab_full, 1);
            mbar_init(&sm.pay_full, 1);
            sm.tile_items[0] = 0; sm.tile_items[1] = 0;
            mbar_init_fence();
        }
    }
    __syncthreads();
    if (FUSED) {
        if (SPEC || !COMPOSE_LEAN) coder_main<SPEC>(sm, f, descs, n_frames, hints, dyns, dyn_tab, payload, out, sizes);
        else coder_lean(sm, f, descs, n_frames, hints, dyns, dyn_tab, payload, out, sizes);
        return;
    }
    /* the register file is re-divided between the groups (whole warpgroups of 128 threads each): the packers give up
       what the coders need (NP * PACKER_REGS + NC * CODER_REGS = NT * the launch allocation) */
    if (tid >= NC) {
#if COMPOSE_PACKER_REGS && !COMPOSE_FUSED
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(COMPOSE_PACKER_REGS));
#endif
        packer_main(sm, tid - NC);
    } else {
#if COMPOSE_PACKER_REGS && !COMPOSE_FUSED
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(COMPOSE_CODER_REGS));
#endif
        if (SPEC) coder_main<true>(sm, f, descs, n_frames, hints, dyns, dyn_tab, payload, out, sizes);
        else coder_lean(sm, f, descs, n_frames, hints, dyns, dyn_tab, payload, out, sizes);
    }
}

/* ------------------------------------------------------------ rasteriser */

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_MAX_REGIONS = H264GPU_MAX_REGIONS;
#ifndef COMPOSE_RASTER_SUB
#define COMPOSE_RASTER_SUB 2
#endif
constexpr int RASTER_SUB = COMPOSE_RASTER_SUB;

/*
 * A warp paints one row at a time: the regions that touch the row are found once (one region per lane, ballot),
 * then every lane carries two neighbouring macroblocks through those regions in list order - a later region simply
 * overwrites an earlier one, which is "later occludes earlier" - and stores them as one 16-byte word.  No division,
 * no per-macroblock walk over regions that do not touch the row.
 */
__global__ void __launch_bounds__(RASTER_THREADS)
raster_regions_kernel(const H264GpuRasterDesc *__restrict__ rdescs, int n_frames,
                      const H264GpuRegion *__restrict__ regions,
                      const H264GpuMbHint *__restrict__ dyn_info,
                      H264GpuMbHint *__restrict__ hints) {
    constexpr int NWR = RASTER_THREADS / 32, NBLK = RASTER_MAX_REGIONS / 32;
    __shared__ uint4 s_reg[RASTER_MAX_REGIONS];
    __shared__ u32 s_act[NWR][NBLK];
    const int f = blockIdx.x;
    if (f >= n_frames) return;
    const H264GpuRasterDesc rd = rdescs[f];
    const int W = rd.mb_w, H = rd.mb_h;
    const int nreg = min((int)rd.region_count, RASTER_MAX_REGIONS);
    const int nblk = (nreg + 31) >> 5;
    for (int r = threadIdx.x; r < nreg; r += RASTER_THREADS)
        s_reg[r] = __ldg((const uint4 *)(regions + rd.region_first) + r);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 *dst = (uint2 *)hints + rd.hint_off;
    const uint2 *dinfo = (const uint2 *)dyn_info + rd.dyn_info_off;
    for (int y = blockIdx.y * NWR + warp; y < H; y += gridDim.y * NWR) {
        for (int k = 0; k < nblk; k++) {                    /* which regions touch this row */
            const int r = 32 * k + lane;
            bool hit = false;
            if (r < nreg) {
                const uint4 q = s_reg[r];
                hit = y >= (int16_t)(q.x >> 16) && y < (int16_t)(q.y >> 16) && (int16_t)(q.y & 0xFFFF) > (int16_t)(q.x & 0xFFFF);
            }
            const u32 m = __ballot_sync(0xFFFFFFFFu, hit);
            if (lane == 0) s_act[warp][k] = m;
        }
        __syncwarp();
        const long long row0 = (long long)y * W;
        for (int xb = 0; xb < W; xb += 64 * RASTER_SUB) {
            /* RASTER_SUB runs of 64 macroblocks per pass (two per lane and run: a warp's stores stay contiguous) */
            uint2 h[RASTER_SUB][2];
#pragma unroll
            for (int u = 0; u < RASTER_SUB; u++) { h[u][0] = make_uint2(0u, 0u); h[u][1] = make_uint2(0u, 0u); }   /* uncovered: (ref 0, mv 0) */
            const int x = xb + 2 * lane;
            for (int k = 0; k < nblk; k++) {
                u32 m = s_act[warp][k];
                while (m) {
                    const uint4 q = s_reg[32 * k + __ffs((int)m) - 1];
                    m &= m - 1u;
                    const int x0 = (int16_t)(q.x & 0xFFFF), x1 = (int16_t)(q.y & 0xFFFF);
                    if (((q.w >> 8) & 0xFFu) == H264GPU_MB_INTER) {
                        const uint2 v = make_uint2(q.z, q.w & 0xFFFFu);
#pragma unroll
                        for (int u = 0; u < RASTER_SUB; u++) {
                            const int xu = x + 64 * u;
                            if (xu >= x0 && xu < x1) h[u][0] = v;
                            if (xu + 1 >= x0 && xu + 1 < x1) h[u][1] = v;
                        }
                    } else {                                 /* per-macroblock hints of a spliced rectangle */
                        const long long base = (long long)(q.w >> 16) + (long long)(y - (int16_t)(q.x >> 16)) * (x1 - x0) - x0;
#pragma unroll
                        for (int u = 0; u < RASTER_SUB; u++) {
                            const int xu = x + 64 * u;
                            if (xu >= x0 && xu < x1) { const uint2 d = __ldg(dinfo + base + xu); h[u][0] = make_uint2(canonical_hint_x(d.x, d.y), canonical_hint_y(d.y)); }
                            if (xu + 1 >= x0 && xu + 1 < x1) { const uint2 d = __ldg(dinfo + base + xu + 1); h[u][1] = make_uint2(canonical_hint_x(d.x, d.y), canonical_hint_y(d.y)); }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RASTER_SUB; u++) {
                const int xu = x + 64 * u;
                const long long g = row0 + xu;
                if (xu + 1 < W && !((rd.hint_off + g) & 1)) *(uint4 *)(dst + g) = make_uint4(h[u][0].x, h[u][0].y, h[u][1].x, h[u][1].y);
                else {
                    if (xu < W) dst[g] = h[u][0];
                    if (xu + 1 < W) dst[g + 1] = h[u][1];
                }
            }
        }
        __syncwarp();                                        /* the row's region masks are replaced next */
    }
}

/* dense maps that did not come from the rasteriser: spliced intra MBs -> (mv 0, ref 0xFF) */
__global__ void __launch_bounds__(256)
canonicalize_hints_kernel(uint2 *__restrict__ hints, unsigned long long n) {
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
        const uint2 h = hints[i];
        if (((h.y >> 8) & 0xFFu) == H264GPU_MB_DYN_INTRA) hints[i] = make_uint2(0u, h.y | KEY_INTRA);
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
    const int smem = (int)sizeof(ComposeSmem);
    /* (per call: function attributes and occupancy belong to the calling thread's current device) */
    cudaError_t e = cudaFuncSetAttribute(compose_frames_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(compose_frames_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = 0;                  /* resident CTAs per SM for this build (registers / shared memory) */
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, compose_frames_kernel<false>, NT, smem) != cudaSuccess || ctas_per_sm < 1)
        ctas_per_sm = 1;
    const int resident = sm_count * ctas_per_sm;      /* persistent grid: a multiple of the SM count */
    int grid = n_frames < resident ? n_frames : resident;
    /* reference-mode frames, then the frames with SPEC flags (that launch returns at once when there are none) */
    compose_frames_kernel<false><<<grid, NT, smem, stream>>>(d_desc, n_frames, d_hints, d_dyn, d_dyn_tab, d_payload, d_out, d_sizes);
    compose_frames_kernel<true><<<grid, NT, smem, stream>>>(d_desc, n_frames, d_hints, d_dyn, d_dyn_tab, d_payload, d_out, d_sizes);
    return cudaGetLastError();
}

cudaError_t launch_canonicalize(H264GpuMbHint *d_hints, size_t n_hints, cudaStream_t stream) {
    if (n_hints == 0) return cudaSuccess;
    size_t blocks = (n_hints + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    canonicalize_hints_kernel<<<(int)blocks, 256, 0, stream>>>((uint2 *)d_hints, (unsigned long long)n_hints);
    return cudaGetLastError();
}

cudaError_t launch_raster(const H264GpuRasterDesc *d_rdesc, int n_frames, int max_mbs,
                          const H264GpuRegion *d_regions, const H264GpuMbHint *d_dyn_info,
                          H264GpuMbHint *d_hints, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    /* rows are dealt out to warps; frames are at most ~max_mbs rows tall, a CTA takes 16 rows per grid row */
    int bx = (max_mbs + 16 * 120 - 1) / (16 * 120);
    if (bx < 1) bx = 1;
    if (bx > 32) bx = 32;
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
