/*
 * h264_writer.h - H.264 syntax writer API, ABI-compatible with the reference's
 * include/h264_writer.h:18-135 (same macros, same public structs, same ten
 * functions).  What changed underneath:
 *
 *   h264_write_scroll_p_frame / h264_write_waypoint_p_frame keep their
 *   signature and side effects (bytes appended to the NALWriter, frame_num++,
 *   waypoint registration) but the macroblock pass, bit packing and emulation
 *   prevention run on the GPU through include/composer_gpu.h.  They return 0
 *   and print a diagnostic when no sm_100 device is available - there is no
 *   CPU fallback.
 *
 *   SPS/PPS generation, I-frame rewriting, slice headers and the waypoint table
 *   stay on the host.
 */
#ifndef H264_WRITER_H
#define H264_WRITER_H

#include <stddef.h>
#include <stdint.h>
#include "bitwriter.h"
#include "nal.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SLICE_TYPE_P     0
#define SLICE_TYPE_I     2
#define SLICE_TYPE_I_ALL 7

#define MV_LIMIT_PX   496     /* vertical MV reach kept below the 512 px decoder limit */
#define MAX_WAYPOINTS 8

typedef struct {
    int offset_px;
    int long_term_idx;
    int valid;
} WaypointInfo;

typedef struct {
    int width, height;        /* pixels */
    int mb_width, mb_height;
    int log2_max_frame_num;
    int pic_order_cnt_type;
    int log2_max_pic_order_cnt_lsb;
    int num_ref_idx_l0_default_minus1;
    int deblocking_filter_control_present_flag;
    int frame_num;
    int idr_pic_id;
    WaypointInfo waypoints[MAX_WAYPOINTS];
    int num_waypoints;
} ComposerConfig;

void   composer_config_init(ComposerConfig *cfg, int width, int height);
void   composer_config_set_sps_params(ComposerConfig *cfg, int log2_max_frame_num,
                                      int pic_order_cnt_type, int log2_max_pic_order_cnt_lsb);
void   composer_config_set_pps_params(ComposerConfig *cfg, int num_ref_idx_l0_default_minus1,
                                      int deblocking_filter_control_present_flag);
size_t h264_generate_sps(uint8_t *rbsp, size_t capacity, int width, int height);
size_t h264_generate_pps(uint8_t *rbsp, size_t capacity);
size_t h264_rewrite_idr_frame(NALWriter *nw, ComposerConfig *write_cfg, ComposerConfig *parse_cfg,
                              const uint8_t *rbsp, size_t rbsp_size);
size_t h264_rewrite_as_non_idr_i_frame(NALWriter *nw, ComposerConfig *write_cfg,
                                       ComposerConfig *parse_cfg, const uint8_t *rbsp,
                                       size_t rbsp_size, int frame_num);
size_t h264_write_scroll_p_frame(NALWriter *nw, ComposerConfig *cfg, int offset_px);
int    h264_needs_waypoint(ComposerConfig *cfg, int offset_px);
size_t h264_write_waypoint_p_frame(NALWriter *nw, ComposerConfig *cfg, int offset_px);

/* ---- additions (not in the reference) ---------------------------------- */

/* Slice-header bits of the next P-frame of this stream, MSB first in hdr[4]
 * (the layout of H264GpuFrameDesc.hdr_bits).  waypoint_frame selects the
 * reference-picture header with MMCO 4+6.  Returns the bit count. */
int    h264_p_slice_header_bits(const ComposerConfig *cfg, int waypoint_frame, uint64_t hdr[4]);

/* The (at most two) motion regions of a legacy scroll frame as
 * {split_row, ref_top, mv_top_px, ref_bottom, mv_bottom_px} (src/h264_writer.c:555-617). */
void   h264_scroll_regions(const ComposerConfig *cfg, int offset_px, int waypoint_frame, int out[5]);

/* The composer's per-frame sequence (src/composer.c:255-264: optional waypoint
 * frame, then the scroll frame) for n offsets in ONE GPU batch.  Returns the
 * number of NAL units appended (waypoint frames included) or -1. */
int    h264_write_scroll_sequence(NALWriter *nw, ComposerConfig *cfg, const int *offsets, int n,
                                  int *waypoints_made);

#ifdef __cplusplus
}
#endif
#endif /* H264_WRITER_H */
