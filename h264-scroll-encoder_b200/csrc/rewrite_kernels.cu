/*
 * rewrite_kernels.cu - the reference's I-frame path on the device.
 *
 * h264_rewrite_idr_frame / h264_rewrite_as_non_idr_i_frame (src/h264_writer.c:242-350): the NAL unit
 *     00 00 00 01 | nal_header | EBSP( new slice header bits | source RBSP bits [start, start + n) | zero padding )
 * i.e. copy_bits (src/h264_writer.c:228-240) over megabytes plus nal_write_unit (src/nal.c:52-84).  The byte string in
 * front of emulation prevention is cut into chunks of one window (32 KB), handed out in launch order by an atomic
 * ticket; a CTA builds its chunk with funnel shifts from cp.async-staged source words, runs the emit stage
 * (emulation prevention + store) on it, and learns the insertions of the chunks before it through a decoupled
 * look-back - one kernel, the source is read once.
 */
#include <cuda_runtime.h>
#include "mb_codec.cuh"
#include "kernel_util.cuh"
#include "emit_stage.cuh"
#include "compose_kernels.cuh"

namespace h264scroll {

namespace {

constexpr int NT = REWRITE_THREADS;      /* threads per CTA */
constexpr int NW = NT / 32;
constexpr int CH = REWRITE_WINDOW_WORDS; /* window: string words per chunk */
constexpr u32 FULL = 0xFFFFFFFFu;
constexpr int EPASS = EMIT_EG * NT;      /* emit: groups per pass */
constexpr int STAGE_WORDS = CH + 32;     /* staged source words of a chunk */

struct __align__(16) RewriteSmem {
    u32 win[CH + 8];                     /* string words of the chunk, MSB first (+ read-ahead pad) */
    u32 stage[STAGE_WORDS];              /* the source words the chunk needs (cp.async) */
    u32 gi[EPASS];                       /* emit: insert mask << 16 | valid bytes << 8 | insert count of each group */
    u32 gpre[EPASS];                     /* emit: bytes inserted before a listed group */
    unsigned short elist[EPASS];         /* emit: pass-local indices of the listed groups */
    unsigned short elist2[EPASS];        /* emit: the listed groups that go out byte-wise */
    u32 n_list, n_list2;
    uint2 scan2[NW + 1];
    u32 zrun[2];                         /* zero bytes closing the bytes emitted so far (rbsp_to_ebsp state) */
    u32 cap;                             /* output slot capacity */
    uint8_t *fout;                       /* output slot */
};

constexpr int RW_CHUNK_BYTES = CH * 4;

__device__ __forceinline__ u32 rewrite_word(const H264GpuRewriteDesc &d, const u64 (&hdr)[4], const u32 *__restrict__ src,
                                            u32 n_src_words, u64 a /* string bit of the word's first bit */) {
    const u64 p0 = 40u + d.hdr_nbits, e = p0 + d.n_bits;
    u32 v = 0u;
    if (a < p0) {
        v = header_bits32(hdr, d.nal_header, (u32)a);
        if (a + 32u > p0) v &= ~(FULL >> (u32)(p0 - a));
    }
    if (a + 32u > p0 && a < e) {
        u32 pv;
        if (a >= p0) pv = pay_ld32(src, n_src_words, (u32)(d.src_start_bit + (a - p0)));
        else pv = pay_ld32(src, n_src_words, (u32)d.src_start_bit) >> (u32)(p0 - a);
        if (a + 32u > e) pv &= ~(FULL >> (u32)(e - a));
        v |= pv;
    }
    return v;
}

static_assert(RW_CHUNK_BYTES / 16 <= EPASS, "a chunk must be one emit pass (the look-back sits inside the pass)");

__global__ void __launch_bounds__(NT, REWRITE_MIN_CTAS)
rewrite_slices_kernel(const H264GpuRewriteDesc *__restrict__ descs, int n_slices, int max_chunks,
                      const uint8_t *__restrict__ src_base, uint8_t *__restrict__ out, u32 *__restrict__ sizes,
                      u32 *__restrict__ state /* [1 + n_slices * max_chunks], zeroed: ticket counter, then the look-back words */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RewriteSmem &sm = *reinterpret_cast<RewriteSmem *>(smem_raw);
    const int tid = threadIdx.x;
    /* chunks are handed out in launch order: whoever holds ticket k knows that every chunk below k is running or done */
    if (tid == 0) sm.n_list = atomicAdd(state, 1u);
    __syncthreads();
    const int ticket = (int)sm.n_list;
    __syncthreads();
    /* tickets walk chunk-major over the slices (chunk c of every slice before chunk c + 1 of any): the CTAs in flight then
       work on a few chunk indices of many slices, so a chunk's predecessors are long done - an inclusive look-back word
       one or two steps back - instead of all being in flight with it */
    const int c = ticket / n_slices, sl = ticket - c * n_slices;
    if (c >= max_chunks) return;
    const H264GpuRewriteDesc d = descs[sl];
    const u64 total_bits = 40u + d.hdr_nbits + d.n_bits;
    const u64 total_bytes = (total_bits + 7u) >> 3;
    const u64 byte0 = (u64)c * RW_CHUNK_BYTES;
    if (byte0 >= total_bytes) return;                        /* whole CTA */
    const int chunk_bytes = (int)min((u64)RW_CHUNK_BYTES, total_bytes - byte0);
    const u64 hdr[4] = {d.hdr_bits[0], d.hdr_bits[1], d.hdr_bits[2], d.hdr_bits[3]};
    const u32 *src = (const u32 *)(src_base + d.src_off);
    const u32 n_src_words = (u32)((d.src_bytes + 3u) >> 2);
    /* ---- the chunk (plus the read-ahead pad of the emit stage) as MSB-first words ----
       The source words the chunk needs are copied to shared memory first (16-byte cp.async, everything in flight at
       once); a word that lies wholly in the copied macroblock data is then
       one funnel shift of two staged words with the same shift for the whole chunk.  Words that touch the header or
       the end of the data go the general way. */
    const int n_words = (chunk_bytes + 3) >> 2;
    const u64 p0 = 40u + d.hdr_nbits, e_bit = p0 + d.n_bits, a0 = 8u * byte0;
    const u64 a_lo = a0 > p0 ? a0 : p0;                      /* first string bit of the chunk that is macroblock data */
    u32 *stage = sm.stage;
    static_assert(STAGE_WORDS >= CH + 16, "the staging buffer must hold a chunk of source words");
    const u64 sbit_lo = d.src_start_bit + (a_lo - p0);       /* source bit of string bit a_lo */
    const u64 src_w_lo = sbit_lo >> 5;                       /* first source word needed */
    const uintptr_t g_addr = ((uintptr_t)(src + src_w_lo)) & ~(uintptr_t)15;          /* 16-byte aligned start of the copy */
    const u32 w_skip = (u32)(((uintptr_t)(src + src_w_lo) - g_addr) >> 2);             /* staged words in front of src_w_lo */
    const bool any_data = a_lo < e_bit && a_lo < a0 + 32u * (u64)n_words;
    if (any_data) {
        const uintptr_t g_end = (uintptr_t)src + (((uintptr_t)d.src_bytes + 3u) & ~(uintptr_t)3) + 16u;   /* readable bytes end here */
        const u32 s_base = smem_u32(stage);
        for (int c16 = tid; c16 < (CH + 16) / 4; c16 += NT) {
            const uintptr_t g = g_addr + 16u * (uintptr_t)c16;
            if (g + 16u <= g_end) cp_async16(s_base + 16u * (u32)c16, (const void *)g);
            else ((uint4 *)stage)[c16] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    cp_async_commit();
    /* ---- zero bytes right in front of the chunk (the state rbsp_to_ebsp would be in, src/nal.c:29-47): one warp looks them
       up while the staging copies are in flight ---- */
    if (tid < 32) {
        u32 run = 0;
        if (c > 0) {
            u64 wi = byte0 / 4u;                             /* words before the chunk, newest first, 32 at a time */
            bool open = true;
            /* a word that lies wholly in the copied data is one funnel shift of two source words, the same shift and the
               same word distance for all of them: inside a long run of zero bytes (a black stripe of an I_PCM atlas is
               kilobytes of them) the walk goes on for many steps, and this is all a step costs */
            const long long delta = (long long)d.src_start_bit - (long long)p0;     /* source bit = string bit + delta */
            const long long dw = delta >> 5;
            const int dsh = (int)(delta & 31);
            while (open && wi > 0) {
                const u64 mine = wi > (u64)tid ? wi - 1u - (u64)tid : ~0ull;
                u32 v = 0xFFFFFFFFu;
                const bool inside = mine != ~0ull && 32u * mine >= p0 && 32u * mine + 32u <= e_bit &&
                                    (u64)((long long)mine + dw) + 1u < (u64)n_src_words;
                if (__all_sync(FULL, inside)) {
                    const u32 *q = src + ((long long)mine + dw);
                    v = fshl(__byte_perm(__ldg(q), 0u, 0x0123u), __byte_perm(__ldg(q + 1), 0u, 0x0123u), dsh);
                } else if (mine != ~0ull) {
                    v = rewrite_word(d, hdr, src, n_src_words, 32u * mine);
                    if (mine == 1u) v |= 0xFF000000u;        /* byte 4 (the NAL header) and the start code never take part */
                    if (mine == 0u) v = 0xFFFFFFFFu;
                }
                const u32 nz = __ballot_sync(FULL, v != 0u);
                if (nz) {
                    const int l = __ffs((int)nz) - 1;        /* nearest word that holds a non-zero byte */
                    const u32 vv = __shfl_sync(FULL, v, l);
                    run += 4u * (u32)l + ((vv & 0xFFu) ? 0u : (vv & 0xFF00u) ? 1u : (vv & 0xFF0000u) ? 2u : 3u);
                    open = false;
                } else { run += 128u; wi = wi > 32u ? wi - 32u : 0u; }
            }
        }
        if (tid == 0) {
            sm.zrun[0] = run; sm.zrun[1] = 0; sm.n_list = 0;
            sm.cap = d.out_cap; sm.fout = out + d.out_off;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    {
        /* string bit a (>= p0) is source bit src_start_bit + a - p0: the word index moves with w, the shift does not.
           Words [w_first, w_last) of the chunk lie wholly inside the copied data */
        const int w_first = a0 >= p0 ? 0 : (int)((p0 - a0 + 31u) >> 5);
        const int w_last = e_bit <= a0 ? 0 : (int)min((u64)n_words, (e_bit - a0) >> 5);
        const u64 sb_first = d.src_start_bit + (a0 + 32u * (u64)w_first - p0);       /* (only meaningful when w_first < w_last) */
        const int sh = (int)(sb_first & 31u);
        const int k_off = (int)((sb_first >> 5) - src_w_lo) + (int)w_skip - w_first;  /* staged word of chunk word w: w + k_off */
        /* four window words per thread and step: five staged words, each byte-swapped once, four funnel shifts, one
           16-byte store; a group that touches the header or the end of the data goes word by word */
        static_assert((CH + 8) % 4 == 0, "the window is built in groups of four words");
        for (int w = 4 * tid; w < CH + 8; w += 4 * NT) {
            if (w >= w_first && w + 4 <= w_last) {
                /* the five staged words sit in two aligned 16-byte chunks (conflict-free loads); which five is the same
                   for every thread: (w + k_off) & 3 does not depend on w */
                const int b = w + k_off;
                const uint4 A = ((const uint4 *)stage)[b >> 2], B = ((const uint4 *)stage)[(b >> 2) + 1];
                u32 s0, s1, s2, s3, s4;
                switch (b & 3) {
                    case 0:  s0 = A.x; s1 = A.y; s2 = A.z; s3 = A.w; s4 = B.x; break;
                    case 1:  s0 = A.y; s1 = A.z; s2 = A.w; s3 = B.x; s4 = B.y; break;
                    case 2:  s0 = A.z; s1 = A.w; s2 = B.x; s3 = B.y; s4 = B.z; break;
                    default: s0 = A.w; s1 = B.x; s2 = B.y; s3 = B.z; s4 = B.w; break;
                }
                s0 = __byte_perm(s0, 0u, 0x0123u); s1 = __byte_perm(s1, 0u, 0x0123u); s2 = __byte_perm(s2, 0u, 0x0123u);
                s3 = __byte_perm(s3, 0u, 0x0123u); s4 = __byte_perm(s4, 0u, 0x0123u);
                *(uint4 *)(sm.win + w) = make_uint4(fshl(s0, s1, sh), fshl(s1, s2, sh), fshl(s2, s3, sh), fshl(s3, s4, sh));
            } else {
#pragma unroll 1
                for (int j = w; j < w + 4; j++) {
                    u32 v = 0u;
                    if (j >= w_first && j < w_last)
                        v = fshl(__byte_perm(stage[j + k_off], 0u, 0x0123u), __byte_perm(stage[j + k_off + 1], 0u, 0x0123u), sh);
                    else if (j < n_words) v = rewrite_word(d, hdr, src, n_src_words, a0 + 32u * (u64)j);
                    sm.win[j] = v;
                }
            }
        }
    }
    __syncthreads();
    volatile u32 *st = state + 1 + (size_t)sl * max_chunks;
    /* (with a look-back the emit stage returns the insertions up to and including this chunk) */
    /* (the insertions before this chunk are not known yet: at most one per two bytes) */
    const u32 ins_incl = emit_window_fits<RewriteSmem, NT>(sm, (u32)byte0, chunk_bytes, (u32)(byte0 / 2u + 1u))
                             ? emit_window_warps<RewriteSmem, NT>(sm, (u32)byte0, chunk_bytes, 0u, 0u, st, c)
                             : emit_window<RewriteSmem, NT>(sm, (u32)byte0, chunk_bytes, 0u, 0u, st, c);
    if (tid == 0 && byte0 + (u64)chunk_bytes >= total_bytes) sizes[sl] = (u32)total_bytes + ins_incl;   /* the slice's last chunk */
}

}  /* anonymous namespace */

cudaError_t launch_rewrite(const H264GpuRewriteDesc *d_desc, int n_slices, int max_chunks, const uint8_t *d_src,
                           uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_state, cudaStream_t stream) {
    if (n_slices <= 0 || max_chunks <= 0) return cudaSuccess;
    const int smem = (int)sizeof(RewriteSmem);
    /* (set on every call: the attribute belongs to the current device's copy of the function) */
    cudaError_t e = cudaFuncSetAttribute(rewrite_slices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)n_slices * (unsigned)max_chunks;
    e = cudaMemsetAsync(d_state, 0, sizeof(uint32_t) * (1 + (size_t)grid), stream);   /* ticket + look-back words */
    if (e == cudaSuccess) e = cudaMemsetAsync(d_sizes, 0, sizeof(uint32_t) * (size_t)n_slices, stream);   /* a slice with more chunks than max_chunks reports 0 */
    if (e != cudaSuccess) return e;
    rewrite_slices_kernel<<<grid, NT, smem, stream>>>(d_desc, n_slices, max_chunks, d_src, d_out, d_sizes, d_state);
    return cudaGetLastError();
}

int rewrite_chunk_bytes() { return RW_CHUNK_BYTES; }

} /* namespace h264scroll */
