/*
 * gpu_session.h - host-side plumbing between the reference-shaped C API
 * (h264_writer.c, composer.c) and the C-ABI GPU entry points of
 * include/composer_gpu.h.  Internal to the library.
 */
#ifndef H264SCROLL_GPU_SESSION_H
#define H264SCROLL_GPU_SESSION_H

#include <stddef.h>
#include <stdint.h>
#include "composer_gpu.h"

typedef struct {
    uint64_t hdr[4];
    int hdr_nbits;
    int nal_header;
    int num_refs;
    int split_row;            /* rows < split_row use the "top" motion */
    int ref_top, mv_top_qpel;
    int ref_bot, mv_bot_qpel;
} LegacyFrame;

/*
 * Compose n legacy scroll frames of one mb_w x mb_h stream on the GPU and
 * append them, in order, at dst (capacity cap).  Returns 0 and the per-frame
 * byte counts in sizes[], or a negative H264GPU_ERR_* code.
 */
int gpu_session_compose_legacy(const LegacyFrame *frames, int n, int mb_w, int mb_h,
                               uint8_t *dst, size_t cap, uint32_t *sizes, size_t *total);

#endif
