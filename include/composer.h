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

#ifdef __cplusplus
}
#endif
#endif /* COMPOSER_H */
