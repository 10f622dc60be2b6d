/*
 * ref_hinted.c - TEST INFRASTRUCTURE ("generalised reference oracle").
 *
 * Compiled only by oracle/Makefile, only where /root/reference exists, into
 * oracle/_ref/libref_hinted.so.  It #includes the reference's own
 * src/h264_writer.c *in place* (nothing is copied into this repository) so
 * that the reference's UNMODIFIED static functions
 *     median3            src/h264_writer.c:362
 *     get_mv_prediction  src/h264_writer.c:369
 *     write_p16x16_mb    src/h264_writer.c:434
 *     copy_bits          src/h264_writer.c:228
 * and its bitwriter.c / nal.c do all the arithmetic.  Only the outer
 * macroblock loop (src/h264_writer.c:595-650) is restated here, reading the
 * per-MB (ref, mv, kind) from a dense hint map instead of deriving it from
 * offset_px.  The reference has no hints / P_Skip / splice of its own, so for
 * those inputs this file is the closest thing to "the reference run here".
 *
 * Never linked into the product library.
 */
#include "h264_writer.c"            /* found through -I$(REF)/src */
#include "composer_gpu.h"           /* hint / flag ABI only */
#include <stdio.h>
#include <time.h>

/* emit the host-written slice header bits, MSB first */
static void put_hdr_bits(BitWriter *bw, const uint64_t *hdr, int nbits) {
    for (int k = 0; k < nbits; k++)
        bitwriter_write_bit(bw, (int)((hdr[k >> 6] >> (63 - (k & 63))) & 1));
}

/*
 * One hinted P-frame -> Annex-B NAL appended at out[0..].  Returns bytes
 * written.  Supports skip OFF / SPEC with the REFERENCE predictor only (the
 * SPEC predictor has no counterpart in the reference).
 */
size_t refhint_compose_frame(const H264GpuMbHint *hints, int mb_w, int mb_h,
                             int num_refs, int flags,
                             const uint64_t *hdr_bits, int hdr_nbits,
                             int nal_header,
                             const uint8_t *payload, size_t payload_bytes,
                             const uint32_t *dyn_bit_off,
                             uint8_t *out, size_t out_cap,
                             uint8_t *rbsp_scratch, size_t rbsp_cap) {
    BitWriter bw;
    bitwriter_init(&bw, rbsp_scratch, rbsp_cap);
    put_hdr_bits(&bw, hdr_bits, hdr_nbits);

    MVInfo *above_row = calloc(mb_w, sizeof(MVInfo));
    MVInfo *current_row = calloc(mb_w, sizeof(MVInfo));
    MVInfo left = {0};
    int skip_count = 0;
    const int skip_spec = (flags & H264GPU_SKIP_SPEC) != 0;

    for (int mb_y = 0; mb_y < mb_h; mb_y++) {
        left.available = 0;
        for (int mb_x = 0; mb_x < mb_w; mb_x++) {
            const H264GpuMbHint *h = &hints[(size_t)mb_y * mb_w + mb_x];
            int ref_idx = h->ref_idx;
            int mvx = h->mv_x_qpel, mvy = h->mv_y_qpel;

            if (h->kind == H264GPU_MB_INTER) {
                int px, py;
                get_mv_prediction(mb_x, mb_y, mb_w, above_row, &left, ref_idx, &px, &py);

                int skipped = 0;
                if (skip_spec && ref_idx == 0) {
                    /* H.264 8.4.1.1: inferred P_Skip motion */
                    int a_av = (mb_x > 0 && left.available);
                    int b_av = (mb_y > 0 && above_row[mb_x].available);
                    int sx, sy;
                    if (!a_av || !b_av ||
                        (left.ref_idx == 0 && left.mv_x == 0 && left.mv_y == 0) ||
                        (above_row[mb_x].ref_idx == 0 && above_row[mb_x].mv_x == 0 &&
                         above_row[mb_x].mv_y == 0)) {
                        sx = 0; sy = 0;
                    } else {
                        sx = px; sy = py;     /* predictor evaluated for ref 0 */
                    }
                    skipped = (mvx == sx && mvy == sy);
                }
                if (skipped) {
                    skip_count++;
                } else {
                    bitwriter_write_ue(&bw, skip_count);
                    skip_count = 0;
                    write_p16x16_mb(&bw, ref_idx, mvx - px, mvy - py, num_refs);
                }
            } else {
                /* spliced pre-encoded macroblock: ue(skip_run) then its bits */
                bitwriter_write_ue(&bw, skip_count);
                skip_count = 0;
                uint32_t b0 = dyn_bit_off[h->dyn_idx], b1 = dyn_bit_off[h->dyn_idx + 1];
                /* the reference's copy_bits walks to start_bit one bit at a time, so hand it the
                 * byte that holds the first bit: same function, same result, no O(offset) walk */
                copy_bits(&bw, payload + (b0 >> 3), payload_bytes - (b0 >> 3), b0 & 7, b1 - b0);
                if (h->kind == H264GPU_MB_DYN_INTRA) { ref_idx = 255; mvx = 0; mvy = 0; }
            }

            current_row[mb_x].mv_x = mvx;
            current_row[mb_x].mv_y = mvy;
            current_row[mb_x].ref_idx = ref_idx;
            current_row[mb_x].available = 1;
            left = current_row[mb_x];
        }
        MVInfo *tmp = above_row; above_row = current_row; current_row = tmp;
    }
    if (skip_count > 0) bitwriter_write_ue(&bw, skip_count);
    free(above_row);
    free(current_row);

    bitwriter_write_trailing_bits(&bw);
    size_t rbsp_size = bitwriter_get_size(&bw);

    NALWriter nw;
    nal_writer_init(&nw, out, out_cap, NULL, 0);
    return nal_write_unit(&nw, (nal_header >> 5) & 3, nal_header & 31,
                          rbsp_scratch, rbsp_size, 1);
}

/* Slice-header bits exactly as the reference writes them (src/h264_writer.c:455-539). */
int refhint_slice_header(ComposerConfig *cfg, int is_waypoint_frame,
                         uint8_t *out_bytes, size_t cap) {
    BitWriter bw;
    bitwriter_init(&bw, out_bytes, cap);
    int frame_num = cfg->frame_num % (1 << cfg->log2_max_frame_num);
    if (is_waypoint_frame)
        h264_write_p_slice_header_waypoint(&bw, cfg, frame_num, frame_num * 2, 1,
                                           2 + cfg->num_waypoints);
    else if (cfg->num_waypoints > 0)
        h264_write_p_slice_header_waypoint(&bw, cfg, frame_num, frame_num * 2, 0, -1);
    else
        h264_write_p_slice_header(&bw, cfg, frame_num, frame_num * 2, 0);
    int nbits = (int)bitwriter_get_bit_position(&bw);
    bitwriter_get_size(&bw);   /* flush the partial byte, left aligned */
    return nbits;
}

/*
 * The composer's per-frame sequence (src/composer.c:255-264) on a bare
 * ComposerConfig prepared the way composer_init + composer_write_header leave
 * it (src/composer.c:192-203; frame_num == 2 after the two I-frames).
 * Records the size of every NAL it appends.  Returns total bytes.
 */
size_t refhint_legacy_stream(int width, int height, int deblock_flag,
                             const int *offsets, int n_offsets,
                             uint8_t *out, size_t out_cap,
                             uint32_t *nal_sizes, int *n_nals, int max_nals) {
    ComposerConfig cfg;
    composer_config_init(&cfg, width, height);
    composer_config_set_sps_params(&cfg, 4, 2, 4);
    composer_config_set_pps_params(&cfg, 1, deblock_flag);
    cfg.frame_num = 2;

    NALWriter nw;
    nal_writer_init(&nw, out, out_cap, NULL, 0);
    int k = 0;
    for (int i = 0; i < n_offsets; i++) {
        if (h264_needs_waypoint(&cfg, offsets[i])) {
            size_t w = h264_write_waypoint_p_frame(&nw, &cfg, offsets[i]);
            if (k < max_nals) nal_sizes[k] = (uint32_t)w;
            k++;
        }
        size_t w = h264_write_scroll_p_frame(&nw, &cfg, offsets[i]);
        if (k < max_nals) nal_sizes[k] = (uint32_t)w;
        k++;
    }
    *n_nals = k;
    return nal_writer_get_size(&nw);
}

/*
 * CPU-baseline timing loops (single thread each; callers run one per core).
 * Return seconds spent composing `reps` passes over the given frames.
 */
double refhint_time_hinted(const H264GpuMbHint *hints, int n_frames, int mb_w, int mb_h,
                           int num_refs, int flags,
                           const uint64_t *hdr_bits, int hdr_nbits,
                           const uint8_t *payload, size_t payload_bytes,
                           const uint32_t *dyn_bit_off, int tab_stride,
                           int reps, uint64_t *bytes_out) {
    size_t cap = 4u << 20;
    uint8_t *out = malloc(cap), *scratch = malloc(cap);
    struct timespec t0, t1;
    uint64_t total = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++)
        for (int f = 0; f < n_frames; f++)
            total += refhint_compose_frame(hints + (size_t)f * mb_w * mb_h, mb_w, mb_h,
                                           num_refs, flags, hdr_bits, hdr_nbits, 0x01,
                                           payload, payload_bytes,
                                           dyn_bit_off ? dyn_bit_off + (size_t)f * tab_stride : NULL,
                                           out, cap, scratch, cap);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(out); free(scratch);
    if (bytes_out) *bytes_out = total;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

double refhint_time_legacy(int width, int height, const int *offsets, int n_offsets,
                           int reps, uint64_t *bytes_out, uint64_t *nals_out) {
    size_t cap = 64u << 20;                       /* src/composer.c:8 */
    uint8_t *out = malloc(cap);
    uint32_t sizes[8];
    int n_nals = 0;
    struct timespec t0, t1;
    uint64_t total = 0, nals = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++) {
        total += refhint_legacy_stream(width, height, 1, offsets, n_offsets, out, cap,
                                       sizes, &n_nals, 8);
        nals += n_nals;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(out);
    if (bytes_out) *bytes_out = total;
    if (nals_out) *nals_out = nals;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
