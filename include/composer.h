/*
 * composer.h - session API, ABI-compatible with the reference's
 * include/composer.h:23-101.  A Composer is caller-allocated, single-owner.
 * composer_write_scroll_frame() now drives the GPU path; everything else
 * (loading the two atlases, header NALs, file output) is host code as before.
 */
#ifndef COMPOSER_H
#define COMPOSER_H

#include <stddef.h>
#include <stdint.h>
#include "h264_writer.h"
#include "nal.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    ComposerConfig cfg;         /* parameters of the stream we write        */
    ComposerConfig parse_cfg;   /* parameters of the atlases' own encoder   */
    uint8_t *ref_a_rbsp;
    size_t ref_a_size;
    uint8_t *ref_b_rbsp;
    size_t ref_b_size;
    uint8_t *orig_sps;
    size_t orig_sps_size;
    uint8_t *orig_pps;
    size_t orig_pps_size;
    NALWriter nw;
    uint8_t *output_buffer;
    size_t output_capacity;
    uint8_t *rbsp_temp;
    size_t rbsp_capacity;
    int frames_written;
} Composer;

int      composer_init(Composer *c, const char *ref_a_path, const char *ref_b_path);  /* 0 / -1 */
int      composer_get_width(Composer *c);
int      composer_get_height(Composer *c);
void     composer_write_header(Composer *c);
void     composer_write_scroll_frame(Composer *c, int offset_px);
size_t   composer_get_output_size(Composer *c);
uint8_t *composer_get_output(Composer *c);
int      composer_write_to_file(Composer *c, const char *path);                       /* 0 / -1 */
void     composer_finish(Composer *c);

/* ---- addition: many frames per launch ----------------------------------
 * Appends the P-frames for offsets[0..n) exactly as n successive
 * composer_write_scroll_frame() calls would (waypoint frames included), but
 * composes them in one GPU batch.  Returns 0, or -1 on a device error. */
int      composer_write_scroll_frames(Composer *c, const int *offsets, int n);

/* ---- addition: hinted frames --------------------------------------------
 *
 * The UI hints of docs/MASTER_DESIGN.md:58-64,105-133: per frame a list of motion-only rectangles and the
 * dynamic rectangle whose macroblocks arrive pre-encoded (:142-171).  The reference has no code for them (its only
 * per-frame input is offset_px, include/composer.h:79); here they ride on top of the scroll frame:
 *
 *   composer_write_hinted_frames(c, frames, n) appends, for every frame, exactly what composer_write_scroll_frame(c,
 *   frames[i].offset_px) would - the waypoint frame first when the offset asks for one (src/composer.c:257-262), with
 *   the same frame_num / long-term-reference bookkeeping - except that the scroll frame's macroblocks under
 *   frames[i].regions carry those regions' motion (later regions occlude earlier ones, the dynamic rectangle occludes
 *   everything, docs/MASTER_DESIGN.md:213-217) and the macroblocks of frames[i].dyn are the caller's pre-encoded bits,
 *   spliced in at their macroblock addresses.  Waypoint frames are plain scroll frames: they become long-term
 *   references and must not hold dynamic content.  All n frames go through the GPU in one batch.
 *   Returns 0, or -1 (invalid hints, device error); on -1 nothing was appended.
 */
typedef struct {
    int16_t mb_x0, mb_y0, mb_x1, mb_y1;   /* half-open macroblock rectangle (clipped to the frame) */
    int16_t mv_x_px, mv_y_px;             /* displacement in whole pixels, as the reference codes them (x4 quarter-pel) */
    uint8_t ref_idx;                      /* 0 = RefA, 1 = RefB, 2 + i = waypoint i of this stream */
} ComposerMotionRegion;

typedef struct {
    int16_t mv_x_qpel, mv_y_qpel;         /* what neighbours see of a pre-encoded macroblock (their MV predictors) */
    uint8_t ref_idx;
    uint8_t intra;                        /* 1: intra coded (neighbours see ref -1, mv 0; src/h264_writer.c:636-640) */
} ComposerMbInfo;

typedef struct {
    int16_t mb_x0, mb_y0, mb_x1, mb_y1;   /* where the macroblocks go, row-major inside the rectangle: n = w * h of them */
    const uint8_t  *payload;              /* their bits, MSB first */
    size_t          payload_bytes;
    const uint32_t *mb_bit_off;           /* n + 1 offsets: macroblock j owns payload bits [off[j], off[j + 1]) */
    const ComposerMbInfo *mb_info;        /* n records */
} ComposerDynRegion;

typedef struct {
    int offset_px;                        /* the scroll position underneath (two bands over RefA / RefB / waypoints) */
    const ComposerMotionRegion *regions;  /* motion-only rectangles over it */
    int n_regions;
    const ComposerDynRegion *dyn;         /* the dynamic rectangle on top, or NULL */
} ComposerFrameHints;

int      composer_write_hinted_frames(Composer *c, const ComposerFrameHints *frames, int n);

#ifdef __cplusplus
}
#endif
#endif /* COMPOSER_H */
