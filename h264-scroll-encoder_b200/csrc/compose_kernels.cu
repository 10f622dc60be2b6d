/*
 * compose_kernels.cu - sm_100a kernels of the P-slice composer.
 *
 * compose_frames_kernel: persistent grid (SMs x 2 CTAs of 512 threads, grid-stride over frames); one CTA composes one
 * frame at a time, all of its warps moving through the stages together (two CTAs per SM overlap each other's barriers;
 * a warp-specialised coder / packer split of the same work was measured at 1.2-1.9x the time, profiles/r2_kernel_experiments.txt).
 * A frame is cut into segments of 8 consecutive macroblocks of one row; a thread owns one segment, a tile is up to 512
 * consecutive segments (4096 macroblocks).
 *
 *   stage    the hints of the tile plus the row above it come global -> shared through the TMA unit: the dense maps are one
 *            2-D tensor of 128-byte rows, a tile is three boxes of 96 rows issued by ONE thread (cp.async.bulk.tensor.2d,
 *            SWIZZLE_128B = the layout the readers use, completion on an mbarrier); frames that do not start on a 128-byte
 *            row or are not a multiple of 8 macroblocks wide take the same tile with 16-byte cp.async.  The copy of the next
 *            tile - of the next frame when this one ends - is in flight while the window is filled / emitted.  The first two
 *            tiles of the next frame's map are sent to L2 a frame ahead, further tiles one tile ahead
 *            (cp.async.bulk.prefetch.L2), and the frame's pre-encoded payload comes to shared memory with ONE bulk copy
 *            (cp.async.bulk.shared.global, completion counted on an mbarrier)
 *   decide   per MB: A/B/C(D) neighbours from shared memory, packed int16 predictor, P_Skip test, mvd, closed-form
 *            Exp-Golomb -> a <= 32-bit code + length in registers.  Where a whole warp sits in uniform motion the
 *            constant code is taken instead; frames without a spliced region try one whole-warp "rows of one motion" vote
 *            first (replaces src/h264_writer.c:595-641, :369-453, src/bitwriter.c:50-101)
 *   skiprun  SPEC mode only: block prefix-max of "last coded MB" -> mb_skip_run prefixes
 *   scan     block prefix sum of the threads' bit counts -> slice bit offsets
 *   write    a thread whose 8 codes are all short concatenates them branch-free into a 2/3/4/8-word integer (picked per
 *            warp), shifts it to its slice bit offset and stores whole words into the shared-memory window (32 KB);
 *            partial first/last words of neighbouring threads are merged by shuffle
 *   queue    everything else - spliced MBs, long codes, separately prefixed codes, the slice trailer - becomes 16-byte
 *            work items that are spread over all lanes of the CTA: one lane splices one pre-encoded MB (copy_bits,
 *            src/h264_writer.c:228) with funnel shifts, whole words as plain stores, edge words with atomicOr
 *   emit     (emit_stage.cuh) emulation prevention (src/nal.c:24-50) on 16-byte groups: the rare groups - two neighbouring
 *            zero bytes inside, the frame's tail - are compacted into a list first and worked through densely; groups
 *            without insertions leave as re-aligned 32/128-bit words, a window without any straight from where it lies;
 *            the window is cleared with one st.bulk
 *
 * The Annex-B start code, the NAL header byte and the host-written slice header are the first bits of the same bit
 * string, so the window image is the output image.
 */
#include <cuda_runtime.h>
#include <cuda.h>
#include <string.h>
#define COMPOSE_TRACE_TU 1
#include "mb_codec.cuh"
#include "kernel_util.cuh"
#include "emit_stage.cuh"
#include "compose_kernels.cuh"

namespace h264scroll {

constexpr int NT = COMPOSE_THREADS;      /* threads per CTA */
constexpr int MPT = 8;                   /* macroblocks per segment / thread */
constexpr int NW = NT / 32;
constexpr int CH = COMPOSE_WINDOW_WORDS; /* window: slice words staged per pass */
constexpr int LCAP = COMPOSE_ITEM_SLOTS; /* queue items per round */
constexpr int HBS = COMPOSE_HINT_SLOTS;  /* 64-byte segment slots per staged hint tile */
constexpr int TBOX_ROWS = 96;            /* TMA box: 96 rows of 128 bytes (12 KB) of the dense map seen as a [rows][32 x u32] tensor: three boxes per tile */
constexpr int TBOX_BYTES = TBOX_ROWS * 128;
constexpr u32 PF_AHEAD = 2u * HBS * 64u;   /* bytes of the next frame's dense map sent to L2 a frame ahead: two tiles */
static_assert((HBS * 64) % TBOX_BYTES == 0, "the hint tile buffer is a whole number of TMA boxes");
constexpr u32 FULL = 0xFFFFFFFFu;
constexpr int PAYBUF16 = COMPOSE_PAYBUF16;
constexpr int EPASS = EMIT_EG * NT;      /* emit: groups per pass */
static_assert(LCAP * 16 >= EPASS * 4, "sm.gi aliases sm.items");

#ifndef COMPOSE_ROW_VOTE
#define COMPOSE_ROW_VOTE 1               /* 1: frames without a spliced region try one whole-warp "rows of one motion" vote first (decide_segment) */
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

struct __align__(1024) ComposeSmem {
    uint4 hbuf[HBS * 4];                 /* the staged hint tile (swizzled 16-byte chunks; 1024-byte aligned: the TMA unit's 128-byte
                                            swizzle works on shared-memory address bits 4-9) */
    u32 win[CH + 8];                     /* slice words [wbase, wbase + CH) of the frame, MSB first (+ read-ahead pad); zero where unwritten */
    union {
        uint4 items[LCAP];               /* work queue (compose stages) */
        u32 gi[EPASS];                   /* emit: insert mask << 16 | valid bytes << 8 | insert count of each group */
    };
    union {
        struct {
            u32 gpre[EPASS];             /* emit: bytes inserted before a listed group */
            unsigned short elist[EPASS]; /* emit: pass-local indices of the listed groups */
            unsigned short elist2[EPASS];/* emit: the listed groups that go out byte-wise */
        };
        uint4 paybuf[PAYBUF16];          /* the frame's pre-encoded payload, staged with cp.async at frame start and dead
                                            by the time the frame's emit stage runs (an emit in the middle of a frame
                                            invalidates it: pay_staged = 0, the splice reads global memory again) */
    };
    u64 pay_full;                        /* mbarrier: the frame's payload has landed (bulk copy) */
    u64 hints_full;                      /* mbarrier: the hint tile has landed (TMA tensor loads) */
    const u32 *pay_smem;                 /* payload word i of the frame = pay_smem[i] for i < pay_staged */
    u32 pay_staged;
    u32 n_list, n_list2;
    u32 scan[2][NW + 1];                 /* double buffered: one barrier per block scan */
    uint2 scan2[NW + 1];
    u32 zrun[2];                         /* zero bytes closing the bytes emitted so far (rbsp_to_ebsp state), double buffered */
    u32 n_items[2];                      /* queued items, double buffered per tile */
    /* per-frame constants only the queue and emit stages need: kept out of the register file */
    const u32 *tab;                      /* the frame's payload bit-offset table */
    const u32 *pay_words;                /* the frame's payload */
    u32 n_pay_words;
    u32 cap;                             /* output slot capacity */
    uint8_t *fout;                       /* output slot */
    TRACE_FIELDS
};

/* geometry of a frame as the tile loop needs it */
struct FrameGeo {
    const uint2 *fh;      /* the frame's dense hint map */
    int W, H;
    int tpr;              /* segments (threads) per row */
    int n_seg;            /* tpr * H */
    int seg_max;          /* segments per tile */
    u32 tpr_magic;        /* s / tpr == umulhi(s, tpr_magic) for s < 2^25 (tpr > 1) */
    int tma_seg0;         /* >= 0: the frame's tiles come by TMA; 64-byte segment index of its first hint in the map (even); -1: cp.async */
};

/* s / tpr == umulhi(s, magic[tpr]) for s < 2^25: magic[tpr] = floor((2^32 - 1) / tpr) + 1, tpr = segments per row <= (MAX_MB_W + 7) / 8 */
constexpr int TPR_MAX = (COMPOSE_MAX_MB_W + MPT - 1) / MPT;
struct TprMagic {
    u32 v[TPR_MAX + 1];
    constexpr TprMagic() : v() { v[0] = 0u; for (int t = 1; t <= TPR_MAX; t++) v[t] = 0xFFFFFFFFu / (u32)t + 1u; }
};
__constant__ TprMagic c_tpr_magic = TprMagic();

__device__ __forceinline__ FrameGeo load_geo(const H264GpuFrameDesc *__restrict__ descs, int f,
                                             const H264GpuMbHint *__restrict__ hints, unsigned long long tma_hints = 0) {
    const uint4 *dq = (const uint4 *)(descs + f);
    const uint4 d2 = __ldg(dq + 2), d3 = __ldg(dq + 3);
    FrameGeo g;
    g.fh = (const uint2 *)hints + ((u64)d2.x | ((u64)d2.y << 32));
    g.W = (int)(d3.z & 0xFFFFu);
    g.H = (int)(d3.z >> 16);
    g.tpr = (g.W + MPT - 1) / MPT;
    g.n_seg = g.tpr * g.H;
    g.seg_max = min(NT, HBS - g.tpr - 2);                    /* (one slot in reserve: a TMA tile starts on a 128-byte row) */
    g.tpr_magic = c_tpr_magic.v[g.tpr];                      /* 0xFFFFFFFF / tpr + 1 from constant memory: the division is ~25 instructions per warp and frame */
    /* TMA: the map's rows are 128 bytes = 16 hints; the frame must start on a row, tile the map with whole segments, and
       end inside the rows the tensor map covers (tma_hints: 0 when the launch has no tensor map) */
    const u64 h0 = (u64)d2.x | ((u64)d2.y << 32);
    g.tma_seg0 = ((g.W & 7) == 0 && (h0 & 15u) == 0 && h0 + (u64)g.W * (u64)g.H <= tma_hints && (h0 >> 3) < 0x7FFF0000ull) ? (int)(h0 >> 3) : -1;
    return g;
}

__device__ __forceinline__ int seg_row(const FrameGeo &g, int s) {
    return g.tpr == 1 ? s : (int)__umulhi((u32)s, g.tpr_magic);
}

/*
 * Stage the hints of segments [s0 - tpr - 1, s0 + n) into buf: slot q holds segment s0 - tpr - 1 + q
 * (8 hints, the tail of a row's last segment stays unwritten), chunks swizzled.  Segments before the
 * frame's first are filled with the "unavailable" hint.
 */
/* frames whose width is not a multiple of 8 macroblocks: chunk by chunk, with its row and column (rare: out of line) */
__device__ __noinline__ void stage_hints_ragged(uint4 *buf, const uint2 *fh, int W, int tpr, u32 tpr_magic, int n_seg, int s_base, int n_chunks) {
    const u32 buf_s = smem_u32(buf);
    for (int c = threadIdx.x; c < n_chunks; c += NT) {
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

/* TMA: 2-D tensor load of one box (TBOX_ROWS rows of 128 bytes from row `row` on) into 128-byte-swizzled shared memory */
__device__ __forceinline__ void tma_load_box(void *dst_smem, const CUtensorMap *tmap, int row, u64 *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}

/*
 * The same tile by TMA (frames whose geometry allows it, FrameGeo::tma_seg0 >= 0): the dense map is a [rows][128 bytes]
 * tensor, the tile a run of rows fetched as boxes of TBOX_ROWS rows by ONE thread; SWIZZLE_128B puts 16-byte chunk c of
 * the tile at chunk c ^ ((c >> 3) & 7) - hint_swizzle, the layout the readers use.  The tile starts on a row: slot q
 * holds segment (seg0 + s0 - tpr - 1 - shift) + q with shift = parity of the first segment wanted.  Rows in front of the
 * frame arrive with whatever the map holds there (zeros outside the tensor); the readers' side overwrites those slots.
 */
__device__ __forceinline__ void stage_hints_tma(ComposeSmem &sm, const CUtensorMap *tmap, const FrameGeo &g, int s0, int n) {
    if (threadIdx.x == 0) {
        const int abs_base = g.tma_seg0 + s0 - g.tpr - 1;
        const int shift = abs_base & 1;
        const int n_rows = (n + g.tpr + 1 + shift + 1) >> 1;
        const int boxes = (n_rows + TBOX_ROWS - 1) / TBOX_ROWS;
        /* the buffer was read (and, in a frame's first tile, partly written) through the generic proxy: order that first */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&sm.hints_full, (u32)(boxes * TBOX_BYTES));
        for (int i = 0; i < boxes; i++)
            tma_load_box((char *)sm.hbuf + i * TBOX_BYTES, tmap, ((abs_base - shift) >> 1) + i * TBOX_ROWS, &sm.hints_full);
        /* the tile behind this one on its way to L2 (tiles past the two that the frame-ahead prefetch covers) */
        const int sn = s0 + n;
        if (64u * (u32)sn >= PF_AHEAD && sn < g.n_seg)
            bulk_prefetch_l2(g.fh + 8 * (long long)sn, 64u * (u32)min(g.seg_max, g.n_seg - sn));
    }
}

__device__ __forceinline__ void stage_hints(uint4 *buf, const FrameGeo &g, int s0, int n) {
    const int s_base = s0 - g.tpr - 1;
    const int n_chunks = 4 * (n + g.tpr + 1);
    if ((g.W & 7) == 0) {
        /* segments tile the dense map exactly: chunk c of the tile is chunk 4 * s_base + c of the map.  A thread's
           chunks are NT apart, which leaves the swizzle term unchanged: one address each side, then fixed strides */
        const int tid = threadIdx.x;
        const u32 dst = smem_u32(buf) + 16u * hint_swizzle((u32)tid);
        uint4 *dst_g = buf + hint_swizzle((u32)tid);
        const uint4 *src = (const uint4 *)g.fh + (4 * (long long)s_base + tid);
        const int c_lo = -4 * s_base;                        /* first chunk inside the frame */
        const u32 c_span = (u32)(4 * g.n_seg);               /* chunks of the frame */
        static_assert(NT % 64 == 0, "the swizzle term must not change between a thread's chunks");
#pragma unroll
        for (int i = 0; i < (HBS * 4 + NT - 1) / NT; i++) {
            const int c = tid + i * NT;
            if (c < n_chunks) {
                if ((u32)(c - c_lo) < c_span) cp_async16(dst + 16u * (u32)(i * NT), src + i * NT);
                else if (c < c_lo) dst_g[i * NT] = make_uint4(0u, HINT_Y_UNAVAIL, 0u, HINT_Y_UNAVAIL);
            }
        }
    } else stage_hints_ragged(buf, g.fh, g.W, g.tpr, g.tpr_magic, g.n_seg, s_base, n_chunks);
}

/* hint i (0..7) of slot q / 16-byte chunk of a staged tile */
__device__ __forceinline__ uint2 staged_hint(const uint4 *buf, int q, int i) {
    const u32 pc = hint_swizzle((u32)(4 * q + (i >> 1)));
    return ((const uint2 *)buf)[2u * pc + (u32)(i & 1)];
}
__device__ __forceinline__ uint4 staged_chunk(const uint4 *buf, int q, int sub) {
    return buf[hint_swizzle((u32)(4 * q + sub))];
}

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
 * The queue.  A thread whose macroblocks cannot all go through the concatenation path leaves ONE record - its bit offset,
 * kinds, codes and lengths as they stand in its registers (plus its own hints when a code is longer than 32 bits and has to be
 * rebuilt: the hint tile may be on its way out by the time the record is read), eight 16-byte words - and the consumer side
 * spreads the records' macroblocks over all lanes of the CTA: lane i takes macroblock i & 7 of record i >> 3, finds its bit offset and prefix
 * from the record, and splices / ORs it in.  (Writing one item per macroblock was the longest stage of a tile for the few
 * warps that had such threads - nine lanes in 32, eight items each - while the others waited.)
 */
constexpr int REC16 = 8;                  /* 16-byte words per record: header, ref keys, 8 codes, 8 lengths, 8 motion vectors */
constexpr int RCAP = LCAP / REC16;        /* records per round */

__device__ __noinline__ void consume_round(ComposeSmem &sm, u32 n_rec, u32 wbase, int num_refs, int skip_spec) {
    const u32 *pw = sm.pay_words, *ps = sm.pay_smem;
    const u32 n_words = sm.n_pay_words, staged = sm.pay_staged;
    for (u32 i = threadIdx.x; i < 8u * n_rec; i += NT) {
        const uint4 *rec = &sm.items[REC16 * (i >> 3)];
        const int k = (int)(i & 7u);
        const uint4 h0 = rec[0], la = rec[4], lb = rec[5];
        const u32 lens[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        u32 bit = h0.x, len = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j < k) bit += lens[j];
            if (j == k) len = lens[j];
        }
        if (!len) continue;
        const u32 kinds = h0.y, kd = (kinds >> (4 * k)) & 0xFu;
        u32 pre = 1u;
        if (skip_spec) {
            /* mb_skip_run + 1: distance to the coded macroblock before this one (the thread's own, or the one handed in) */
            const u32 below = k ? kinds & ((1u << (4 * k)) - 1u) : 0u;
            const int prev = below ? (int)h0.z + (int)(flo32(below) >> 2) : (int)h0.w;
            pre = (u32)((int)h0.z + k - prev);
        }
        const u32 *cw = (const u32 *)(rec + 2) + k;          /* the code word as it lies in the record */
        uint2 h = make_uint2(0u, 0u);
        if (kd == MBK_LONG) {                                /* its own hint, left in the record by the thread */
            h.x = ((const u32 *)(rec + 6))[k];
            h.y = (((const u32 *)(rec + 1))[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        }
        const Item it = make_item(kd, *cw, len, bit, pre, h.x, h.y);
        /* a spliced macroblock whose payload words (and the read-ahead word) are all staged reads shared memory */
        const bool in_smem = (it.x >> 30) == IT_DYN && ((it.z + it.w + 31u) >> 5) + 1u < staged;
        /* (every staged word can be read: past the payload's end they are zero; a short code is read where it lies) */
        consume_item(it, cw - 2, sm.win, wbase, (u32)CH, num_refs, in_smem ? ps : pw, in_smem ? staged : n_words);
    }
}

/*
 * The 8 macroblocks of a segment -> codes, lengths, kinds (4 bits each).  GEN = 0 is the reference mode
 * (skip OFF, reference predictor) with everything mode dependent folded away; GEN = 1 takes the flags.
 * A macroblock in uniform motion - every available neighbour carries its own hint - takes the constant
 * zero-mvd code; the vote keeps the warp on that path as long as no lane needs the predictor
 * (spliced macroblocks do not: their length comes from the offset table).
 */
template <int GEN, int DUMMY>
__device__ __forceinline__ u32 decide_segment(const uint2 (&own)[MPT + 1], const uint2 (&abv)[MPT + 2], int nmb, bool first, bool row0,
                                              int num_refs, const u32 *__restrict__ tab, u32 (&code)[MPT], u32 (&len)[MPT],
                                              bool rows, int skip_spec_in = 0, int pred_spec_in = 0) {
    const int skip_spec = GEN ? skip_spec_in : 0, pred_spec = GEN ? pred_spec_in : 0;
    u32 kinds = 0;
#if COMPOSE_ROW_VOTE
    if (rows) {
        /* frames without a spliced region (the reference's own scroll frames, per-row motion): ONE vote - every segment of the
           warp whole, its own hints all alike and the hints above it all alike.  Macroblocks 1..7 then share their inputs
           (own, A = own, B, C = B; at a row's end C is D = the same hint): one code for them, macroblock 0 (its A is the
           segment before, or nothing) by itself */
        u32 d = own[1].y & 0xFF00u;
#pragma unroll
        for (int k = 2; k <= MPT; k++) d |= (own[k].x ^ own[1].x) | (own[k].y ^ own[1].y);
#pragma unroll
        for (int k = 2; k <= MPT + 1; k++) d |= (abv[k].x ^ abv[1].x) | (abv[k].y ^ abv[1].y);
        if (__all_sync(FULL, nmb == 0 || (nmb == MPT && d == 0u))) {
            u32 c0 = 0u, l0 = 0u, c1 = 0u, l1 = 0u, k0 = MBK_NONE, k1 = MBK_NONE;
            if (nmb) {
                k0 = decide_mb(own[1].x, own[1].y, own[0].x, own[0].y, abv[1].x, abv[1].y, abv[2].x, abv[2].y, num_refs, pred_spec, skip_spec, c0, l0);
                k1 = decide_mb(own[2].x, own[2].y, own[1].x, own[1].y, abv[2].x, abv[2].y, abv[3].x, abv[3].y, num_refs, pred_spec, skip_spec, c1, l1);
            }
            code[0] = c0; len[0] = l0;
#pragma unroll
            for (int k = 1; k < MPT; k++) { code[k] = c1; len[k] = l1; }
            return k0 | (k1 * 0x11111110u);
        }
    }
#endif
    /* spliced macroblocks: their payload ranges come from the offset table in global memory - all loads of the segment
       are issued here, in front of the votes, so that they overlap instead of queueing one round trip behind the other */
    u32 any_dyn = 0;
#pragma unroll
    for (int k = 0; k < MPT; k++) {
        code[k] = 0u; len[k] = 0u;
        any_dyn |= own[1 + k].y;
    }
    if (__any_sync(FULL, (any_dyn & 0xFF00u) != 0u)) {           /* (most warps have no spliced macroblock at all) */
#pragma unroll
        for (int k = 0; k < MPT; k++) {
            const u32 hy = own[1 + k].y;
            if (k < nmb && (hy & 0xFF00u) != 0u) {
                const u32 di = hy >> 16;
                code[k] = __ldg(tab + di);
                len[k] = __ldg(tab + di + 1);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MPT; k++) {
        const uint2 h = own[1 + k], a = own[k], b = abv[k + 1], c = abv[k + 2];
        const bool present = k < nmb;
        const bool dyn = (h.y & 0xFF00u) != 0u;
        bool uni;
        if (skip_spec) {
            /* P_Skip looks at A and B themselves: all three neighbours must carry the MB's own hint */
            uni = ((h.x ^ a.x) | (h.y ^ a.y) | (h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
        } else {
            /* B == C == the MB's own hint: two refs match, and both medians of (a, h, h) are h whatever A is.  (The top row -
               A alone decides, src/h264_writer.c:407-417 - has no shortcut of its own: its B and C are the "unavailable" hint,
               the test fails, and the one warp that holds the row takes the predictor once per segment and the "same as the
               macroblock before" path for the rest; a second test here cost every warp of every tile three instructions
               per macroblock) */
            uni = ((h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
        }
        u32 kd = MBK_NONE;
        if (__all_sync(FULL, uni || dyn || !present)) {
            if (present && !dyn) kd = decide_mb_uniform(h.y, num_refs, skip_spec, code[k], len[k]);
        } else {
            /* somebody needs the predictor.  The code is a function of (own hint, A, B, C) alone: along a region's
               horizontal edge all four repeat from macroblock to macroblock, and the answer of the macroblock before
               is taken over (a whole row of such macroblocks costs the predictor once per segment, not eight times) */
            bool same = false;
            if (k > 0)
                same = ((h.x ^ a.x) | (h.y ^ a.y) | (a.x ^ own[k - 1].x) | (a.y ^ own[k - 1].y) |
                        (b.x ^ abv[k].x) | (b.y ^ abv[k].y) | (c.x ^ b.x) | (c.y ^ b.y)) == 0u;
            if (k > 0 && __all_sync(FULL, uni || same || dyn || !present)) {
                if (present && !dyn) {
                    if (same) { kd = (kinds >> (4 * (k - 1))) & 0xFu; code[k] = code[k - 1]; len[k] = len[k - 1]; }
                    else kd = decide_mb_uniform(h.y, num_refs, skip_spec, code[k], len[k]);
                }
            } else if (present && !dyn) {
                kd = decide_mb(h.x, h.y, a.x, a.y, b.x, b.y, c.x, c.y, num_refs, pred_spec, skip_spec, code[k], len[k]);
            }
        }
        if (present && dyn) {
            len[k] = len[k] - code[k] + (skip_spec ? 0u : 1u);
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
    /* (a batch of one mode - the usual case - answers with the first candidate: one broadcast load instead of 32 scattered ones) */
    if (f < n_frames && frame_is_spec(descs, f) == SPEC) return f;
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

template <bool SPEC>
__global__ void __launch_bounds__(NT, COMPOSE_MIN_CTAS)
compose_frames_kernel(const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                      const H264GpuMbHint *__restrict__ hints,
                      const H264GpuDynDesc *__restrict__ dyns,
                      const u32 *__restrict__ dyn_tab,
                      const uint8_t *__restrict__ payload,
                      uint8_t *__restrict__ out, u32 *__restrict__ sizes,
                      const __grid_constant__ CUtensorMap hint_map, unsigned long long tma_hints) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    ComposeSmem &sm = *reinterpret_cast<ComposeSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;

    int f = next_own_frame<SPEC>(descs, blockIdx.x, n_frames, (int)gridDim.x);
    if (f >= n_frames) return;

    {   /* the window is zero wherever nothing was written: cleared once here, then every flush clears what it used */
        uint4 *w4 = (uint4 *)sm.win;
        for (int i = tid; i < (CH + 8) / 4; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) {
            sm.n_items[0] = 0; sm.n_items[1] = 0; sm.n_list = 0;
#if COMPOSE_TRACE
            for (int i = 0; i < 32; i++) sm.tr_n[i] = 0;
            sm.tr_on = 0;
#endif
#if COMPOSE_BULK_OPS
            mbar_init(&sm.pay_full, 1);
            mbar_init(&sm.hints_full, 1);
            mbar_init_fence();
#endif
        }
    }
    __syncthreads();                   /* (the mbarriers exist before anybody waits on them) */
    u32 hpar = 0;                      /* phase of sm.hints_full the next TMA-staged tile completes */
    u32 spar = 0;                      /* which sm.scan buffer the next block scan uses */
#if COMPOSE_BULK_OPS
    u32 pay_par = 0;                   /* phase of sm.pay_full the next payload copy completes */
    bool pay_pending = false;          /* this frame's payload copy has not been waited for yet */
#endif
    u32 ipar = 0;                      /* which sm.n_items counter the current tile uses */
#if COMPOSE_TRACE
    int trace_frame = 0;
#endif

#if COMPOSE_TIMING
    long long stage_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, stage_t = clock64();
#endif
    FrameGeo geo = load_geo(descs, f, hints, tma_hints);
    if (geo.tma_seg0 >= 0) stage_hints_tma(sm, &hint_map, geo, 0, min(geo.seg_max, geo.n_seg));
    else stage_hints(sm.hbuf, geo, 0, min(geo.seg_max, geo.n_seg));
    cp_async_commit();

    while (f < n_frames) {
        /* ---- frame setup: the descriptor is read by every thread (one broadcast line) ------ */
        const uint4 *dq = (const uint4 *)(descs + f);
        const uint4 d2 = __ldg(dq + 2), d3 = __ldg(dq + 3);
        const u64 out_off = (u64)d2.z | ((u64)d2.w << 32);
        const u32 cap = d3.x, dyn_id = d3.y;
        const int W = geo.W;
        const u32 hdr_total = 40u + (d3.w & 0xFFu);
        const int num_refs = (int)((d3.w >> 8) & 0xFFu);
        const int skip_spec = SPEC ? ((d3.w >> 16) & H264GPU_SKIP_SPEC) != 0 : 0;
        const int pred_spec = SPEC ? ((d3.w >> 16) & H264GPU_PRED_SPEC) != 0 : 0;
        const u32 nal_header = d3.w >> 24;
        const u32 *tab = nullptr;
        {
            const u32 *pay_words = nullptr;
            u32 n_pay_words = 0;
            if (dyn_id != H264GPU_NO_DYN) {
                const uint4 *yq = (const uint4 *)(dyns + dyn_id);
                const uint4 y0 = __ldg(yq), y1 = __ldg(yq + 1);
                tab = dyn_tab + ((u64)y0.z | ((u64)y0.w << 32));
                pay_words = (const u32 *)(payload + ((u64)y0.x | ((u64)y0.y << 32)));
                n_pay_words = (y1.y + 3u) >> 2;
            }
            /* the payload goes to shared memory (whole 16-byte chunks of it; what does not fit, and the odd bytes at its
               end, stay in global memory and are only sent on their way to L2): the splice then reads it at shared-memory
               latency.  Its cp.async group is committed behind the group of this frame's first hint tile, so the tile
               loop waits for "all but the last group" until the queue stage needs the payload */
            u32 staged = 0;
            const u32 *pay_smem = nullptr;
            if (pay_words != nullptr && PAYBUF16 > 0) {
                const uintptr_t a0 = (uintptr_t)pay_words & ~(uintptr_t)15;
                const u32 w_skip = (u32)(((uintptr_t)pay_words - a0) >> 2);
                const u32 n16 = min((w_skip + n_pay_words) >> 2, (u32)PAYBUF16);
#if COMPOSE_BULK_OPS
                /* one bulk copy through the TMA unit, counted on an mbarrier; the buffer was last written by the previous
                   frame's emit stage (generic proxy), every thread of which is past its barriers */
                if (tid == 0 && n16) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(&sm.pay_full, 16u * n16);
                    bulk_g2s(sm.paybuf, (const void *)a0, 16u * n16, &sm.pay_full);
                }
                pay_pending = n16 != 0u;
#else
                const u32 dst = smem_u32(sm.paybuf);
                for (u32 c = (u32)tid; c < n16; c += NT) cp_async16(dst + 16u * c, (const void *)(a0 + 16u * (uintptr_t)c));
#endif
                staged = 4u * n16 > w_skip ? 4u * n16 - w_skip : 0u;
                if (n16 + 2u <= (u32)PAYBUF16) {
                    /* the words behind the last whole chunk (up to three) and zero words to read ahead into go in by hand:
                       the macroblocks at the payload's end are spliced from shared memory like all the others */
                    const u32 wi = 4u * n16 + (u32)tid;
                    if (tid < 8) ((u32 *)sm.paybuf)[wi] = wi < w_skip + n_pay_words ? __ldg((const u32 *)a0 + wi) : 0u;
                    staged = 4u * n16 + 8u - w_skip;
                }
                pay_smem = (const u32 *)sm.paybuf + w_skip;
                for (u32 o = 16u * n16 + 128u * (u32)tid; o < 4u * (w_skip + n_pay_words); o += 128u * NT) prefetch_l2((const uint8_t *)a0 + o);
            }
            cp_async_commit();
            if (tid == 0) {                                    /* read again only after this frame's first barrier */
                sm.zrun[0] = 0; sm.zrun[1] = 0;
                sm.tab = tab; sm.pay_words = pay_words; sm.n_pay_words = n_pay_words;
                sm.pay_smem = pay_smem; sm.pay_staged = staged;
                sm.cap = cap; sm.fout = out + out_off;
#if COMPOSE_TRACE
                sm.tr_on = (trace_frame >= 4 && trace_frame < 7) ? 1u : 0u;
#endif
            }
        }
#if COMPOSE_TRACE
        trace_frame++;
#endif
        TRACE(1);
        const int kc = (W - 1) & 7;                            /* position of a row's last MB inside its segment */

        /* frame state every thread tracks identically (all of it follows from block-wide totals) */
        u32 bitpos = hdr_total;        /* slice bits composed so far (incl. the 40-bit Annex-B prefix) */
        int last_coded = -1;           /* raster index of the last non-skipped MB */
        u32 ins_total = 0;             /* emulation-prevention bytes inserted so far */
        u32 zpar = 0;                  /* which sm.zrun slot holds the current state */
        u32 wbase = 0;                 /* frame word held by window word 0; words below it have left for global memory */

        const int f_next = next_own_frame<SPEC>(descs, f + (int)gridDim.x, n_frames, (int)gridDim.x);
        FrameGeo geo_next = geo;
        if (f_next < n_frames) {
            geo_next = load_geo(descs, f_next, hints, tma_hints);
#if COMPOSE_BULK_OPS
            /* the next frame's dense hint map goes to L2 a whole frame ahead (bulk prefetch): its tile copies find it there */
            const u32 n_bytes = (8u * (u32)(geo_next.W * geo_next.H)) & ~15u;
            /* (its first two tiles: what lies behind them follows tile by tile, stage_hints_tma - the whole map of a 4K frame
               a frame ahead, 259 KB x 296 CTAs, does not stay in the 126 MB L2 beside the maps being read: ncu showed 1.7x the
               hint bytes in DRAM reads on config 4) */
            if (tid == 32 && n_bytes) bulk_prefetch_l2(geo_next.fh, n_bytes < PF_AHEAD ? n_bytes : PF_AHEAD);
#endif
        }

        for (int s0 = 0; s0 < geo.n_seg;) {
            const int n = min(geo.seg_max, geo.n_seg - s0);
            const bool last = (s0 + n >= geo.n_seg);
            const bool tile0 = (s0 == 0);
            const uint4 *hb = sm.hbuf;

            /* ---- stage: this tile's hints have landed ---------- */
            STAGE_MARK(7);                                      /* frame setup / loop glue */
            if (tile0) cp_async_wait_but_last(); else cp_async_wait_all();   /* (the payload's group may still be in flight) */
            int shift = 0;                                      /* TMA tiles start on a 128-byte row: slots may sit one further */
            if (geo.tma_seg0 >= 0) {
                mbar_wait(&sm.hints_full, hpar);
                hpar ^= 1u;
                shift = (geo.tma_seg0 + s0 - geo.tpr - 1) & 1;
                if (tile0) {
                    /* the slots in front of the frame's first segment hold the neighbours "above the picture": unavailable */
                    const int n_sent = 4 * (geo.tpr + 1 + shift);
                    for (int c = tid; c < n_sent; c += NT) sm.hbuf[hint_swizzle((u32)c)] = make_uint4(0u, HINT_Y_UNAVAIL, 0u, HINT_Y_UNAVAIL);
                }
            }
            TRACE(10);
            __syncthreads();
            TRACE(11);
            STAGE_MARK(0);                                      /* wait for the hint tile (and for the slowest warp of the tile before) */

            /* ---- load: own segment, the one above, the hints beside them ------------------------ */
            const int s = s0 + tid;
            const bool active = tid < n;
            const int row = seg_row(geo, active ? s : s0), col = (active ? s : s0) - row * geo.tpr;
            const int nmb = active ? min(MPT, W - MPT * col) : 0;
            const int g0 = row * W + MPT * col;                 /* raster index of the segment's first MB */
            const bool first = col == 0, lastseg = col == geo.tpr - 1;
            /* slots of the own segment and of the one above.  (A thread past the tile's end reads thread 0's: what it holds is
               never looked at - nmb = 0 keeps it out of every vote and every store - and filling its 38 registers with the
               "unavailable" hint instead cost every thread 38 predicated moves per tile) */
            const int tq = active ? tid : 0;
            const int qo = tq + geo.tpr + 1 + shift, qa = tq + 1 + shift;

            uint2 own[MPT + 1];                                 /* own[0] = left neighbour of the first MB */
            uint2 abv[MPT + 2];                                 /* abv[j] = hint above-left of MB j */
            {
                /* chunk 4 q + k of the tile lives at (4 q + k) ^ ((q >> 1) & 7) = ((4 q) ^ ((q >> 1) & 7)) ^ k: one swizzle term per
                   slot, the chunk number folded in as a constant */
                const u32 xo = (4u * (u32)qo) ^ (((u32)qo >> 1) & 7u), xa = (4u * (u32)qa) ^ (((u32)qa >> 1) & 7u);
#pragma unroll
                for (int k = 0; k < MPT / 2; k++) {
                    const uint4 v = hb[xo ^ (u32)k], u = hb[xa ^ (u32)k];
                    own[1 + 2 * k] = make_uint2(v.x, v.y); own[2 + 2 * k] = make_uint2(v.z, v.w);
                    abv[1 + 2 * k] = make_uint2(u.x, u.y); abv[2 + 2 * k] = make_uint2(u.z, u.w);
                }
                own[0] = staged_hint(hb, qo - 1, 7);
                abv[0] = staged_hint(hb, qa - 1, 7);
                abv[MPT + 1] = staged_hint(hb, qa + 1, 0);
                if (first) { own[0] = make_uint2(0u, HINT_Y_UNAVAIL); abv[0] = make_uint2(0u, HINT_Y_UNAVAIL); }
                /* C of the row's last MB is D (src/h264_writer.c:391-399).  (The frame-wide test outside: widths that are a multiple
                   of 8 macroblocks - nearly all - branch over the chain for the other seven positions) */
                if (kc == 7) { if (lastseg) abv[9] = abv[7]; }
                else if (lastseg) {
#pragma unroll
                    for (int j = 0; j < MPT - 1; j++)
                        if (kc == j) abv[j + 2] = abv[j];
                }
            }

            /* ---- decide + code ------------------------------------------------ */
            u32 code[MPT], len[MPT];
            u32 kinds;
            if (!SPEC) kinds = decide_segment<0, 0>(own, abv, nmb, first, row == 0, num_refs, tab, code, len, tab == nullptr);
            else kinds = decide_segment<1, 1>(own, abv, nmb, first, row == 0, num_refs, tab, code, len, tab == nullptr, skip_spec, pred_spec);

            STAGE_MARK(1);                                      /* load + decide */
            /* ---- mb_skip_run prefixes (SPEC skip; with skip OFF the constant bit is already in) ---- */
            int last_coded_total = last_coded;
            int prev0 = 0;
            if (skip_spec) {
                int lc = -1;
#pragma unroll
                for (int k = 0; k < MPT; k++)
                    if ((kinds >> (4 * k)) & 0xFu) lc = g0 + k;
                int blk_max;
                prev0 = max(group_excl_scan_max<NW, 0>(lc, (int *)sm.scan[spar], tid >> 5, blk_max), last_coded);
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
            u32 ibase = 0;                                       /* the thread's record in the queue */
            if (cplx) ibase = atomicAdd(&sm.n_items[ipar], 1u);
            const u32 L = Ls + (owns_trailer ? tr_len : 0u);

            /* ---- bit offsets: prefix sum of the threads' bit counts ----------------- */
            u32 total;
            TRACE(12);
            const u32 O = bitpos + group_excl_scan<NW, 0>(L, sm.scan[spar], tid >> 5, total);
            spar ^= 1u;
            const u32 total_items = sm.n_items[ipar];
            TRACE(13);
            STAGE_MARK(2);                                      /* prefixes, bit counts, scan (incl. its barrier) */
            if (tid == 0) sm.n_items[ipar ^ 1u] = 0;               /* the next tile's counter; nobody adds to it before the next barrier */
            const u32 E = O + L;
            const u32 end = bitpos + total;
            const u32 total_bytes = (end + 7u) >> 3;                       /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* frame words [.., w_stop) are complete after this tile */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end; /* last word this tile writes to */
            const u32 warp_end = __shfl_sync(FULL, E, 31);

            /* ---- the tile does not fit behind what the window already holds: send that off first ---- */
            {
                const u32 wf_al = (bitpos >> 5) & ~3u;             /* 16-byte group of the tile's first word */
                if (!tile0 && w_last - wbase >= (u32)CH && wf_al > wbase) {
                    __syncthreads();
                    const u32 src = wf_al - wbase;
#if COMPOSE_BULK_OPS
                    if (pay_pending) { mbar_wait(&sm.pay_full, pay_par); pay_par ^= 1u; pay_pending = false; }
#endif
                    if (tid == 0) sm.pay_staged = 0;                       /* the emit stage reuses the payload buffer */
                    ins_total += emit_window<ComposeSmem, NT>(sm, wbase << 2, (int)(src << 2), ins_total, zpar);
                    zpar ^= 1u;
                    __syncthreads();                                       /* window consumed */
                    if (tid < 4) { const u32 c = sm.win[src + tid]; sm.win[src + tid] = 0u; sm.win[tid] = c; }
#if COMPOSE_BULK_OPS
                    if (tid == 32 && src > 4u) smem_zero_bulk(sm.win + 4, 4u * (src - 4u));   /* groups 1 .. src/4 - 1: one st.bulk */
#else
                    uint4 *w4 = (uint4 *)sm.win;
                    for (u32 i = 1u + (u32)tid; i < src / 4u; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
#endif
                    wbase = wf_al;
                    __syncthreads();
                    STAGE_MARK(5);
                }
            }

            for (;;) {
                /* ---- header: start code, NAL header byte, slice header bits ---------- */
                if (tile0 && tid < 10 && 32u * (u32)tid < hdr_total) {
                    const uint4 h0 = __ldg(dq), h1 = __ldg(dq + 1);
                    const u64 hdr[4] = {(u64)h0.x | ((u64)h0.y << 32), (u64)h0.z | ((u64)h0.w << 32),
                                        (u64)h1.x | ((u64)h1.y << 32), (u64)h1.z | ((u64)h1.w << 32)};
                    u32 v = header_bits32(hdr, nal_header, 32u * (u32)tid);
                    const u32 left = hdr_total - 32u * (u32)tid;
                    if (left < 32u) v &= ~(FULL >> left);
                    if ((u32)tid - wbase < (u32)CH && v) atomicOr(&sm.win[(u32)tid - wbase], v);
                }

                /* ---- write: concatenation path + merge of the partial words between threads ---------- */
                {
                    const u32 Lc = cplx ? 0u : Ls;                         /* bits through the concatenation path */
                    const u32 Ec = O + Lc;
                    u32 H = 0u, T = 0u;
                    bool has_t = false;
                    const u32 lmax = __reduce_max_sync(FULL, Lc);
                    if (lmax <= 64u) { if (Lc) simple_words<2>(code, len, O, Lc, sm.win, wbase, H, T, has_t); }
                    /* (65..96 bits: a segment of 6-bit codes with one long mvd - the macroblock in front of a region edge) */
                    else if (lmax <= 96u) { if (Lc) simple_words<3>(code, len, O, Lc, sm.win, wbase, H, T, has_t); }
                    else if (lmax <= 128u) { if (Lc) simple_words<4>(code, len, O, Lc, sm.win, wbase, H, T, has_t); }
                    else { if (Lc) simple_words<8>(code, len, O, Lc, sm.win, wbase, H, T, has_t); }
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
                            if (lane == 31) atomicOr(&sm.win[tj - wbase], v);   /* the next warp's (or tile's) first head lands here */
                            else sm.win[tj - wbase] = v;
                        }
                        if (lane == 0 && H != 0u && key - wbase < (u32)CH) atomicOr(&sm.win[key - wbase], H);
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
                            if ((warp_end >> 5) == tj) atomicOr(&sm.win[tj - wbase], v);   /* later warps (or the next tile) add to this word */
                            else sm.win[tj - wbase] = v;
                        }
                        if (lane == 0 && X != 0u && key - wbase < (u32)CH) atomicOr(&sm.win[key - wbase], X);
                    }
                }

                TRACE(14);
                STAGE_MARK(3);                                  /* stage next tile, header, concatenation path + merge */
                /* ---- queue: the records of this tile in rounds of RCAP, one lane per macroblock ----------- */
                for (u32 r0 = 0; r0 < total_items; r0 += (u32)RCAP) {
                    if (r0) __syncthreads();                               /* the previous round is consumed */
                    if (cplx && ibase - r0 < (u32)RCAP) {
                        uint4 *rec = &sm.items[REC16 * (ibase - r0)];
                        rec[0] = make_uint4(O, kinds, (u32)g0, (u32)prev0);
                        {
                            const u32 t = kinds ^ (0x11111111u * MBK_LONG);
                            if ((t - 0x11111111u) & ~t & 0x88888888u) {    /* a code that the consumer rebuilds from predictor and hint */
                                u32 hx[MPT], key[2] = {0u, 0u};
#pragma unroll
                                for (int k = 0; k < MPT; k++) {
                                    const uint2 h = staged_hint(hb, tid + geo.tpr + 1 + shift, k);   /* (its own slot: a queueing thread is inside the tile) */
                                    hx[k] = h.x;
                                    key[k >> 2] |= (h.y & 0xFFu) << (8 * (k & 3));
                                }
                                rec[1] = make_uint4(key[0], key[1], 0u, 0u);
                                rec[6] = make_uint4(hx[0], hx[1], hx[2], hx[3]);
                                rec[7] = make_uint4(hx[4], hx[5], hx[6], hx[7]);
                            }
                        }
                        rec[2] = make_uint4(code[0], code[1], code[2], code[3]);
                        rec[3] = make_uint4(code[4], code[5], code[6], code[7]);
                        rec[4] = make_uint4(len[0], len[1], len[2], len[3]);
                        rec[5] = make_uint4(len[4], len[5], len[6], len[7]);
                    }
#if !COMPOSE_BULK_OPS
                    cp_async_wait_all();                                   /* the staged payload has landed */
#else
                    if (pay_pending) { mbar_wait(&sm.pay_full, pay_par); pay_par ^= 1u; pay_pending = false; }
#endif
                    TRACE(15);
                    __syncthreads();
                    TRACE(16);
                    STAGE_MARK(6);                              /* queue: push + wait for the slowest writer */
                    consume_round(sm, min((u32)RCAP, total_items - r0), wbase, num_refs, skip_spec);
                    TRACE(17);
                }
                /* the slice trailer (mb_skip_run of the skipped tail + rbsp_stop_one_bit) goes in behind its thread's macroblocks */
                if (owns_trailer) win_or64(sm.win, wbase, (u32)CH, O + Ls, (u64)tr_code, tr_len);

                STAGE_MARK(4);                                  /* queue */
                const bool fits = w_last - wbase < (u32)CH;                /* the tile's last word lies inside the window */
                /* ---- stage: everybody is past its loads from the hint tile (the last ones fetch the hints of the long codes
                       in the push loop, in front of a barrier): start the copy of the next tile - of the next frame's first
                       when this is the last.  (A tile written in several window passes pushes again: its last pass stages) */
                if (fits) {
                    if (!last || f_next < n_frames) {
                        const FrameGeo &gn = last ? geo_next : geo;
                        const int sn = last ? 0 : s0 + n;
                        if (gn.tma_seg0 >= 0) stage_hints_tma(sm, &hint_map, gn, sn, min(gn.seg_max, gn.n_seg - sn));
                        else stage_hints(sm.hbuf, gn, sn, min(gn.seg_max, gn.n_seg - sn));
                    }
                    cp_async_commit();
                }
                if (!(last || !fits)) break;                               /* keep filling the window with the next tile */

                /* ---- emit: emulation prevention + store of frame words [wbase, e_hi) --------- */
                TRACE(18);
                __syncthreads();
                TRACE(19);
                const u32 e_hi = fits ? w_stop : wbase + CH;
                const u32 chunk_byte0 = wbase << 2;
                const int chunk_bytes = e_hi > wbase ? (int)((fits ? min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0;
#if COMPOSE_BULK_OPS
                if (pay_pending) { mbar_wait(&sm.pay_full, pay_par); pay_par ^= 1u; pay_pending = false; }
#endif
                if (tid == 0) sm.pay_staged = 0;                               /* the emit stage reuses the payload buffer */
                ins_total += emit_window<ComposeSmem, NT>(sm, chunk_byte0, chunk_bytes, ins_total, zpar);
                if (chunk_bytes > 0) zpar ^= 1u;
                TRACE(20);
                __syncthreads();
                TRACE(21);                                                     /* window fully consumed */
                {   /* clear the words this window used */
                    const u32 used = min((u32)CH, w_last - wbase + 1u);
#if COMPOSE_BULK_OPS
                    if (tid == 0) smem_zero_bulk(sm.win, 4u * ((used + 3u) & ~3u));       /* one st.bulk (UMEMSETS) */
#else
                    uint4 *w4 = (uint4 *)sm.win;
                    for (u32 i = (u32)tid; i < (used + 3u) / 4u; i += NT) w4[i] = make_uint4(0u, 0u, 0u, 0u);
#endif
                }
                STAGE_MARK(5);                                  /* emit + window clear */
                if (fits) break;                                               /* (frame done) */
                wbase += CH;                                                   /* a full window has left; the tile goes on in the next one */
                __syncthreads();
            }

            bitpos = end;
            last_coded = last_coded_total;
            if (last && tid == 0) sizes[f] = total_bytes + ins_total;
            ipar ^= 1u;
            s0 += n;
        }
        f = f_next;
        geo = geo_next;
    }
    cp_async_wait_all();
#if COMPOSE_TIMING
    if (tid == 0)
        for (int i = 0; i < 8; i++) atomicAdd(&g_stage_cycles[i], (unsigned long long)stage_acc[i]);
#endif
}

#if COMPOSE_TRACE
cudaError_t read_trace(unsigned long long *h) { return cudaMemcpyFromSymbol(h, g_trace, sizeof(unsigned long long) * TRACE_WARPS * TRACE_EVENTS); }
#endif
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

/*
 * The dense hint maps as a 2-D tensor for the TMA unit: rows of 128 bytes (16 hints, 32 x u32), boxes of TBOX_ROWS rows,
 * 128-byte swizzle.  Only the whole rows of the n_hints the caller vouches for are covered.  Returns the hints covered
 * (0: no tensor map - the kernel stages every tile with cp.async).
 */
static unsigned long long make_hint_map(CUtensorMap *map, const H264GpuMbHint *d_hints, size_t n_hints) {
    memset(map, 0, sizeof *map);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static bool looked = false;
    if (!looked) {                                       /* (the driver entry point is process-wide; a benign race at worst) */
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = (EncodeFn)fn;
        else cudaGetLastError();
        looked = true;
    }
    const cuuint64_t rows = (cuuint64_t)(n_hints / 16);
    if (!encode || rows == 0 || ((uintptr_t)d_hints & 127u)) return 0;
    const cuuint64_t dims[2] = {32, rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)TBOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)d_hints, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        memset(map, 0, sizeof *map);
        return 0;
    }
    return (unsigned long long)rows * 16ull;
}

cudaError_t launch_compose(const H264GpuFrameDesc *d_desc, int n_frames,
                           const H264GpuMbHint *d_hints, size_t n_hints, const H264GpuDynDesc *d_dyn,
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
#ifdef COMPOSE_FORCE_CTAS_PER_SM              /* experiments: fewer CTAs per SM than fit */
    ctas_per_sm = COMPOSE_FORCE_CTAS_PER_SM;
#endif
    const int resident = sm_count * ctas_per_sm;      /* persistent grid: a multiple of the SM count */
    int grid = n_frames < resident ? n_frames : resident;
    CUtensorMap hint_map;
    const unsigned long long tma_hints = COMPOSE_BULK_OPS ? make_hint_map(&hint_map, d_hints, n_hints) : (memset(&hint_map, 0, sizeof hint_map), 0ull);
    /* reference-mode frames, then the frames with SPEC flags (that launch returns at once when there are none) */
    compose_frames_kernel<false><<<grid, NT, smem, stream>>>(d_desc, n_frames, d_hints, d_dyn, d_dyn_tab, d_payload, d_out, d_sizes, hint_map, tma_hints);
    compose_frames_kernel<true><<<grid, NT, smem, stream>>>(d_desc, n_frames, d_hints, d_dyn, d_dyn_tab, d_payload, d_out, d_sizes, hint_map, tma_hints);
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
