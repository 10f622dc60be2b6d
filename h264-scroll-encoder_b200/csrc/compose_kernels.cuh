/* compose_kernels.cuh - launch interface between capi.cu and compose_kernels.cu */
#ifndef H264SCROLL_COMPOSE_KERNELS_CUH
#define H264SCROLL_COMPOSE_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "composer_gpu.h"

#ifndef COMPOSE_THREADS
#define COMPOSE_THREADS      512     /* threads per CTA: a tile is 4096 macroblocks.  Few large CTAs beat many small ones here:
                                        fewer code regions in flight per SM (the kernel is sensitive to instruction fetch) */
#endif
#ifndef COMPOSE_MIN_CTAS
#define COMPOSE_MIN_CTAS     2       /* resident CTAs per SM the register budget is sized for */
#endif
#ifndef COMPOSE_WINDOW_WORDS
#define COMPOSE_WINDOW_WORDS 8192    /* slice words staged in shared memory per pass (multiple of 4) */
#endif
#ifndef COMPOSE_ITEM_SLOTS
#define COMPOSE_ITEM_SLOTS   1024    /* work-queue items per round */
#endif
#ifndef COMPOSE_HINT_SLOTS
#define COMPOSE_HINT_SLOTS   (COMPOSE_THREADS + 64)   /* 64-byte segment slots per staged hint tile (a whole number of 4 KB TMA boxes):
                                                         a full tile for frames up to 496 MBs wide, at least 2 segments per tile at the widest */
#endif
#ifndef COMPOSE_PAYBUF16
#define COMPOSE_PAYBUF16     1536    /* 16-byte chunks of a frame's payload staged in shared memory (26 KB: a 25x25-MB region at ~300 bits per MB) */
#endif
#ifndef COMPOSE_BULK_OPS
#define COMPOSE_BULK_OPS     1       /* sm_100 bulk operations: payload staged by one cp.async.bulk (TMA unit) on an mbarrier, next frame's hints
                                        sent to L2 by cp.async.bulk.prefetch, window cleared by st.bulk; 0: cp.async / store loops */
#endif
/* slice rewrite: 256-thread CTAs on 16 KB chunks, five per SM (measured: 512 x 2 on 32 KB 0.192 ms, 256 x 4 0.171, 256 x 5
   0.157, 128 x 8..10 on 8 KB 0.160-0.164: the chunk's load -> build -> look-back -> store chain wants many CTAs in flight) */
#ifndef REWRITE_THREADS
#define REWRITE_THREADS      256
#endif
#ifndef REWRITE_MIN_CTAS
#define REWRITE_MIN_CTAS     5
#endif
#ifndef REWRITE_WINDOW_WORDS
#define REWRITE_WINDOW_WORDS 4096    /* string words per rewrite chunk (16 KB = 4 groups of 16 bytes per thread) */
#endif
#define COMPOSE_MAX_MB_W  H264GPU_MAX_MB_W

namespace h264scroll {

/* n_hints = hints the caller vouches for behind d_hints (0: unknown - no TMA tensor map, every tile by cp.async) */
cudaError_t launch_compose(const H264GpuFrameDesc *d_desc, int n_frames,
                           const H264GpuMbHint *d_hints, size_t n_hints, const H264GpuDynDesc *d_dyn,
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
