/* compose_kernels.cuh - launch interface between capi.cu and compose_kernels.cu */
#ifndef H264SCROLL_COMPOSE_KERNELS_CUH
#define H264SCROLL_COMPOSE_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "composer_gpu.h"

#ifndef COMPOSE_FUSED
#define COMPOSE_FUSED        1       /* 1: every thread codes a tile, then packs the window (phase-synchronous CTA); 0: warp-specialised coder /
                                        packer groups over two windows */
#endif
#ifndef COMPOSE_LEAN
#define COMPOSE_LEAN         1       /* reference-mode frames through the rolled coder (coder_lean.cuh); 0: through the general one */
#endif
#ifndef COMPOSE_THREADS
#define COMPOSE_THREADS      512     /* threads per CTA: coder warps + packer warps */
#endif
#ifndef COMPOSE_CODER_THREADS
#define COMPOSE_CODER_THREADS 384    /* coder threads: a tile is 3072 macroblocks (multiple of 64); the rest are packers */
#endif
#ifndef COMPOSE_MIN_CTAS
#define COMPOSE_MIN_CTAS     2       /* resident CTAs per SM the register budget is sized for */
#endif
#ifndef COMPOSE_WINDOW_WORDS
#define COMPOSE_WINDOW_WORDS (COMPOSE_FUSED ? 8192 : 4096)   /* slice words per window buffer (multiple of 4) */
#endif
#ifndef COMPOSE_ITEM_SLOTS
#define COMPOSE_ITEM_SLOTS   (COMPOSE_FUSED ? 1024 : 512)  /* work-queue items per window hand-over */
#endif
#ifndef COMPOSE_HINT_SLOTS
#define COMPOSE_HINT_SLOTS   ((COMPOSE_FUSED ? COMPOSE_THREADS : COMPOSE_CODER_THREADS) + 34)   /* 64-byte segment slots per staged hint tile: a full tile for frames
                                                               up to 264 MBs wide, at least 2 segments per tile at the widest */
#endif
#ifndef COMPOSE_TAB_WORDS
#define COMPOSE_TAB_WORDS    1024    /* entries of a frame's payload bit-offset table staged in shared memory (larger tables are read from global memory) */
#endif
#ifndef COMPOSE_PAYBUF16
#define COMPOSE_PAYBUF16     1408    /* 16-byte chunks of a frame's payload staged in shared memory (24 KB: a 25x25-MB region at ~300 bits per MB) */
#endif
#ifndef COMPOSE_PACKER_REGS
#define COMPOSE_PACKER_REGS  40      /* setmaxnreg: registers per packer thread (0: leave the launch allocation alone) ... */
#define COMPOSE_CODER_REGS   72      /* ... and per coder thread: 128 * 40 + 384 * 72 = 512 * 64 */
#endif
#ifndef COMPOSE_BULK_CLEAR
#define COMPOSE_BULK_CLEAR   1       /* the packers clear a window with one st.bulk instead of a store loop */
#endif
#ifndef REWRITE_THREADS
#define REWRITE_THREADS      512
#endif
#ifndef REWRITE_MIN_CTAS
#define REWRITE_MIN_CTAS     2
#endif
#ifndef REWRITE_WINDOW_WORDS
#define REWRITE_WINDOW_WORDS 8192    /* string words per rewrite chunk (32 KB) */
#endif
#define COMPOSE_MAX_MB_W  H264GPU_MAX_MB_W

namespace h264scroll {

cudaError_t launch_compose(const H264GpuFrameDesc *d_desc, int n_frames,
                           const H264GpuMbHint *d_hints, const H264GpuDynDesc *d_dyn,
                           const uint32_t *d_dyn_tab, const uint8_t *d_payload,
                           uint8_t *d_out, uint32_t *d_sizes, int sm_count, cudaStream_t stream);

cudaError_t launch_canonicalize(H264GpuMbHint *d_hints, size_t n_hints, cudaStream_t stream);

cudaError_t launch_rewrite(const H264GpuRewriteDesc *d_desc, int n_slices, int max_chunks, const uint8_t *d_src,
                           uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_counts, cudaStream_t stream);
int rewrite_chunk_bytes();

cudaError_t launch_unescape(const H264GpuUnescapeDesc *d_desc, int n_nals, int max_chunks, const uint8_t *d_src,
                            uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_state, cudaStream_t stream);
int unescape_chunk_bytes();

cudaError_t launch_parse_annexb(const uint8_t *d_data, size_t n_bytes, H264GpuNalUnit *d_units, uint32_t max_nals,
                                uint32_t *d_scratch, cudaStream_t stream);
size_t parse_annexb_scratch_words(size_t n_bytes, uint32_t max_nals);

cudaError_t launch_raster(const H264GpuRasterDesc *d_rdesc, int n_frames, int max_mbs,
                          const H264GpuRegion *d_regions, const H264GpuMbHint *d_dyn_info,
                          H264GpuMbHint *d_hints, cudaStream_t stream);

cudaError_t launch_pack(const H264GpuFrameDesc *d_desc, const uint32_t *d_sizes, int n_frames,
                        const uint8_t *d_slots, uint8_t *d_packed, unsigned long long *d_offsets,
                        cudaStream_t stream);

}
#endif
