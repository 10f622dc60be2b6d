/*
 * composer.c - session object behind include/composer.h (reference:
 * src/composer.c).  Loads the two atlases, emits the stream header on the host
 * and dispatches every P-frame to the GPU path.  Messages on stdout/stderr are
 * the reference's.
 */
#include "composer.h"
#include "nal_parser.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OUTPUT_BUFFER_SIZE (64u * 1024u * 1024u)   /* src/composer.c:8 */
#define RBSP_BUFFER_SIZE   (4u * 1024u * 1024u)    /* src/composer.c:9 */

typedef struct {
    uint8_t *sps, *pps, *idr;
    size_t sps_size, pps_size, idr_size;
    int width, height;
    int log2_max_frame_num, poc_type, log2_max_poc_lsb;
    int num_ref_idx_l0, deblock_flag;
} Atlas;

static void atlas_free(Atlas *a) {
    free(a->sps); free(a->pps); free(a->idr);
    memset(a, 0, sizeof *a);
}

static uint8_t *dup_bytes(const uint8_t *p, size_t n) {
    uint8_t *d = malloc(n ? n : 1);
    if (d) memcpy(d, p, n);
    return d;
}

/* first SPS, PPS and IDR of an Annex-B file (src/composer.c:45-125) */
static int atlas_load(const char *path, Atlas *a) {
    memset(a, 0, sizeof *a);
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Error: Cannot open %s\n", path); return -1; }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *data = malloc(len > 0 ? (size_t)len : 1);
    if (!data || fread(data, 1, (size_t)len, f) != (size_t)len) {
        fprintf(stderr, "Error: Failed to read %s\n", path);
        free(data); fclose(f);
        return -1;
    }
    fclose(f);

    uint8_t *tmp = malloc((size_t)len ? (size_t)len : 1);
    NALParser ps;
    NALUnit u;
    int rc = 0;
    nal_parser_init(&ps, data, (size_t)len);
    while (rc == 0 && nal_parser_next(&ps, &u)) {
        if (u.nal_unit_type == NAL_TYPE_SPS && !a->sps) {
            size_t n = ebsp_to_rbsp(tmp, u.data, u.size);
            if (parse_sps(tmp, n, &a->width, &a->height, &a->log2_max_frame_num, &a->poc_type,
                          &a->log2_max_poc_lsb) < 0) {
                fprintf(stderr, "Error: Failed to parse SPS\n");
                rc = -1;
            } else { a->sps = dup_bytes(tmp, n); a->sps_size = n; }
        } else if (u.nal_unit_type == NAL_TYPE_PPS && !a->pps) {
            size_t n = ebsp_to_rbsp(tmp, u.data, u.size);
            if (parse_pps(tmp, n, &a->num_ref_idx_l0, &a->deblock_flag) < 0) {
                fprintf(stderr, "Error: Failed to parse PPS\n");
                rc = -1;
            } else { a->pps = dup_bytes(tmp, n); a->pps_size = n; }
        } else if (u.nal_unit_type == NAL_TYPE_IDR && !a->idr) {
            size_t n = ebsp_to_rbsp(tmp, u.data, u.size);
            a->idr = dup_bytes(tmp, n); a->idr_size = n;
        }
    }
    free(tmp);
    free(data);
    if (rc == 0 && (!a->sps || !a->pps || !a->idr)) {
        fprintf(stderr, "Error: Reference file missing SPS/PPS/IDR\n");
        rc = -1;
    }
    if (rc) atlas_free(a);
    return rc;
}

int composer_init(Composer *c, const char *ref_a_path, const char *ref_b_path) {
    memset(c, 0, sizeof *c);
    Atlas a, b;
    int ra = atlas_load(ref_a_path, &a);
    int rb = atlas_load(ref_b_path, &b);
    if (ra || rb) { if (!ra) atlas_free(&a); if (!rb) atlas_free(&b); return -1; }
    if (a.width != b.width || a.height != b.height) {
        fprintf(stderr, "Error: Reference frame dimensions don't match\n");
        fprintf(stderr, "  RefA: %dx%d, RefB: %dx%d\n", a.width, a.height, b.width, b.height);
        atlas_free(&a); atlas_free(&b);
        return -1;
    }
    c->orig_sps = a.sps; c->orig_sps_size = a.sps_size;
    c->orig_pps = a.pps; c->orig_pps_size = a.pps_size;
    c->ref_a_rbsp = a.idr; c->ref_a_size = a.idr_size;
    c->ref_b_rbsp = b.idr; c->ref_b_size = b.idr_size;
    free(b.sps); free(b.pps);

    /* how the atlases were coded ... */
    composer_config_init(&c->parse_cfg, a.width, a.height);
    composer_config_set_sps_params(&c->parse_cfg, a.log2_max_frame_num, a.poc_type, a.log2_max_poc_lsb);
    composer_config_set_pps_params(&c->parse_cfg, a.num_ref_idx_l0, a.deblock_flag);
    /* ... and how we code: frame_num in 4 bits, POC type 2, two default refs (src/composer.c:196-203) */
    composer_config_init(&c->cfg, a.width, a.height);
    composer_config_set_sps_params(&c->cfg, 4, 2, 4);
    composer_config_set_pps_params(&c->cfg, 1, a.deblock_flag);

    c->output_capacity = OUTPUT_BUFFER_SIZE;
    c->output_buffer = malloc(c->output_capacity);
    c->rbsp_capacity = RBSP_BUFFER_SIZE;
    c->rbsp_temp = malloc(c->rbsp_capacity);
    if (!c->output_buffer || !c->rbsp_temp) {
        fprintf(stderr, "Error: Failed to allocate output buffers\n");
        return -1;
    }
    nal_writer_init(&c->nw, c->output_buffer, c->output_capacity, c->rbsp_temp, c->rbsp_capacity);
    printf("Composer initialized: %dx%d\n", a.width, a.height);
    return 0;
}

int composer_get_width(Composer *c) { return c->cfg.width; }
int composer_get_height(Composer *c) { return c->cfg.height; }

/* SPS, PPS, RefA as long-term IDR, RefB as long-term I (src/composer.c:232-253) */
void composer_write_header(Composer *c) {
    size_t n = h264_generate_sps(c->rbsp_temp, c->rbsp_capacity, c->cfg.width, c->cfg.height);
    nal_write_unit(&c->nw, NAL_REF_IDC_HIGHEST, NAL_TYPE_SPS, c->rbsp_temp, n, 1);
    n = h264_generate_pps(c->rbsp_temp, c->rbsp_capacity);
    nal_write_unit(&c->nw, NAL_REF_IDC_HIGHEST, NAL_TYPE_PPS, c->rbsp_temp, n, 1);
    h264_rewrite_idr_frame(&c->nw, &c->cfg, &c->parse_cfg, c->ref_a_rbsp, c->ref_a_size);
    h264_rewrite_as_non_idr_i_frame(&c->nw, &c->cfg, &c->parse_cfg, c->ref_b_rbsp, c->ref_b_size, 1);
    printf("Header written: SPS + PPS + 2 reference frames\n");
}

void composer_write_scroll_frame(Composer *c, int offset_px) {
    if (h264_needs_waypoint(&c->cfg, offset_px)) {
        h264_write_waypoint_p_frame(&c->nw, &c->cfg, offset_px);
        printf("  Waypoint at offset %d\n", offset_px);
    }
    h264_write_scroll_p_frame(&c->nw, &c->cfg, offset_px);
    c->frames_written++;
}

int composer_write_scroll_frames(Composer *c, const int *offsets, int n) {
    int made = 0;
    if (n <= 0) return 0;
    if (h264_write_scroll_sequence(&c->nw, &c->cfg, offsets, n, &made) < 0) return -1;
    c->frames_written += n;
    return 0;
}

int composer_write_hinted_frames(Composer *c, const ComposerFrameHints *frames, int n) {
    int made = 0;
    if (n <= 0) return 0;
    if (!frames || h264_write_hinted_sequence(&c->nw, &c->cfg, frames, n, &made) < 0) return -1;
    c->frames_written += n;
    return 0;
}

size_t composer_get_output_size(Composer *c) { return nal_writer_get_size(&c->nw); }
uint8_t *composer_get_output(Composer *c) { return nal_writer_get_output(&c->nw); }

int composer_write_to_file(Composer *c, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "Error: Cannot create %s\n", path); return -1; }
    size_t n = composer_get_output_size(c);
    if (fwrite(c->output_buffer, 1, n, f) != n) {
        fprintf(stderr, "Error: Failed to write %s\n", path);
        fclose(f);
        return -1;
    }
    fclose(f);
    printf("Written %zu bytes to %s\n", n, path);
    return 0;
}

void composer_finish(Composer *c) {
    free(c->ref_a_rbsp); free(c->ref_b_rbsp);
    free(c->orig_sps); free(c->orig_pps);
    free(c->output_buffer); free(c->rbsp_temp);
    memset(c, 0, sizeof *c);
}
