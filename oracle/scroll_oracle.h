/*
 * scroll_oracle.h - TEST INFRASTRUCTURE.  CPU restatement ("port") of the
 * reference composer's P-slice hot path, used only as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product
 * library never includes, links or calls anything in oracle/.
 *
 * Pinning: tests/test_oracle_pinned.py checks this restatement byte-for-byte
 * against (a) the unmodified reference compiled in place (oracle/_ref, built
 * by oracle/Makefile from /root/reference) and (b) the golden fixtures under
 * tests/golden/ that oracle/gen_golden.py produced from that compiled
 * reference.  Extension inputs the reference cannot express (SPEC predictor)
 * are "parity unpinned": only this file defines them.
 */
#ifndef SCROLL_ORACLE_H
#define SCROLL_ORACLE_H

#include <stdint.h>
#include <stddef.h>
#include "composer_gpu.h"      /* hint / descriptor PODs and flag values */

#define SO_MAX_WAYPOINTS 8     /* include/h264_writer.h:27 */
#define SO_MV_LIMIT_PX 496     /* include/h264_writer.h:24 */

/* Per-stream host state the reference keeps in ComposerConfig
 * (include/h264_writer.h:37-59). */
typedef struct {
    int width, height, mb_w, mb_h;
    int log2_max_frame_num, poc_type, log2_max_poc_lsb;
    int deblock_present;
    int frame_num;
    int n_wp;
    int wp_offset[SO_MAX_WAYPOINTS];
    int wp_lt_idx[SO_MAX_WAYPOINTS];
    int wp_valid[SO_MAX_WAYPOINTS];
} SoStream;

void   so_stream_init(SoStream *st, int width, int height, int deblock_present);

/* P-slice header bits (src/h264_writer.c:455-539) into hdr[4], MSB first.
 * waypoint_frame: 0 = scroll frame, 1 = waypoint frame.  Returns bit count. */
int    so_slice_header(const SoStream *st, int waypoint_frame, uint64_t hdr[4]);

/* Region derivation of src/h264_writer.c:555-617 (scroll) / :689-729 (waypoint):
 * fills a dense hint map. */
void   so_legacy_hints(const SoStream *st, int offset_px, int waypoint_frame,
                       H264GpuMbHint *hints);

int    so_needs_waypoint(const SoStream *st, int offset_px);      /* :666-676 */
void   so_register_waypoint(SoStream *st, int offset_px);         /* :771-777 */

/* One hinted frame -> Annex-B NAL.  Returns bytes written (0 on overflow). */
size_t so_compose_frame(const H264GpuMbHint *hints, int mb_w, int mb_h,
                        int num_refs, int flags,
                        const uint64_t *hdr_bits, int hdr_nbits, int nal_header,
                        const uint8_t *payload, size_t payload_bytes,
                        const uint32_t *dyn_bit_off,
                        uint8_t *out, size_t out_cap);

/* Per-MB intermediate values for stage-by-stage kernel checks.
 * bitlen[i] = bits MB i appends (incl. its mb_skip_run code), 0 for skipped. */
void   so_frame_stages(const H264GpuMbHint *hints, int mb_w, int mb_h,
                       int num_refs, int flags, const uint32_t *dyn_bit_off,
                       int16_t *pred_x, int16_t *pred_y, uint32_t *bitlen);

/* The composer's per-frame sequence (src/composer.c:255-264) over a list of
 * offsets, starting from the state composer_write_header leaves. */
size_t so_legacy_stream(int width, int height, int deblock_present,
                        const int *offsets, int n_offsets,
                        uint8_t *out, size_t out_cap,
                        uint32_t *nal_sizes, int *n_nals, int max_nals);

/* src/nal.c:24-50 */
size_t so_rbsp_to_ebsp(uint8_t *ebsp, size_t cap, const uint8_t *rbsp, size_t n);

/* timing loop for bench.py's cpu_baseline ("port"); seconds for reps passes */
double so_time_hinted(const H264GpuMbHint *hints, int n_frames, int mb_w, int mb_h,
                      int num_refs, int flags, const uint64_t *hdr_bits, int hdr_nbits,
                      const uint8_t *payload, size_t payload_bytes,
                      const uint32_t *dyn_bit_off, int tab_stride,
                      int reps, uint64_t *bytes_out);

#endif
