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

/* what rides on top of a legacy frame (include/composer.h); NULL members = nothing */
typedef struct {
    const void *regions;      /* ComposerMotionRegion[n_regions] */
    int n_regions;
    const void *dyn;          /* ComposerDynRegion or NULL */
} HintOverlay;

/*
 * The same for frames with UI hints on top (overlay[i] may be NULL): region lists, per-macroblock neighbour info,
 * offset tables and payload go through h264gpu_compose_regions_host in one pipelined batch.
 */
int gpu_session_compose_hinted(const LegacyFrame *frames, const HintOverlay *overlay, int n, int mb_w, int mb_h,
                               uint8_t *dst, size_t cap, uint32_t *sizes, size_t *total);


/* h264_writer.c: the composer's per-frame sequence with hints on top (frames = ComposerFrameHints[n], include/composer.h).
   Returns the NAL units appended or -1; on -1 neither the writer nor cfg has changed. */
int h264_write_hinted_sequence(void *nal_writer, void *composer_config, const void *frames, int n, int *waypoints_made);

#endif
