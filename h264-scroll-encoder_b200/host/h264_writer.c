/*
 * h264_writer.c - host side of include/h264_writer.h.
 *
 * Host work that stays host work (north star (c)): parameter sets, I-frame
 * slice-header rewriting, P-slice headers, frame_num and the waypoint /
 * long-term-reference table.  The two P-frame writers keep the reference's
 * signatures and side effects (src/h264_writer.c:541-664, :678-782) but hand
 * the macroblock pass to the GPU via gpu_session.c.
 */
#include "h264_writer.h"
#include "gpu_session.h"
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------ configuration */

void composer_config_init(ComposerConfig *cfg, int width, int height) {
    memset(cfg, 0, sizeof *cfg);
    cfg->width = width;
    cfg->height = height;
    cfg->mb_width = width / 16;             /* truncating, like the reference (:17-18) */
    cfg->mb_height = height / 16;
    cfg->log2_max_frame_num = 4;
    cfg->pic_order_cnt_type = 2;
    cfg->log2_max_pic_order_cnt_lsb = 4;
    cfg->num_ref_idx_l0_default_minus1 = 1;
    cfg->deblocking_filter_control_present_flag = 1;
}

void composer_config_set_sps_params(ComposerConfig *cfg, int log2_max_frame_num,
                                    int pic_order_cnt_type, int log2_max_pic_order_cnt_lsb) {
    cfg->log2_max_frame_num = log2_max_frame_num;
    cfg->pic_order_cnt_type = pic_order_cnt_type;
    cfg->log2_max_pic_order_cnt_lsb = log2_max_pic_order_cnt_lsb;
}

void composer_config_set_pps_params(ComposerConfig *cfg, int num_ref_idx_l0_default_minus1,
                                    int deblocking_filter_control_present_flag) {
    cfg->num_ref_idx_l0_default_minus1 = num_ref_idx_l0_default_minus1;
    cfg->deblocking_filter_control_present_flag = deblocking_filter_control_present_flag;
}

/* ------------------------------------------------------------ parameter sets */

/* Baseline SPS (field order of src/h264_writer.c:49-100) */
size_t h264_generate_sps(uint8_t *rbsp, size_t capacity, int width, int height) {
    BitWriter bw;
    bitwriter_init(&bw, rbsp, capacity);
    bitwriter_write_bits(&bw, (66u << 16) | (0xC0u << 8) | 40u, 24);  /* profile, constraints, level 4.0 */
    bitwriter_write_ue(&bw, 0);                       /* sps id */
    bitwriter_write_ue(&bw, 0);                       /* log2_max_frame_num - 4 */
    bitwriter_write_ue(&bw, 2);                       /* pic_order_cnt_type */
    bitwriter_write_ue(&bw, 2 + MAX_WAYPOINTS);       /* max_num_ref_frames */
    bitwriter_write_bit(&bw, 0);                      /* gaps in frame_num */
    bitwriter_write_ue(&bw, (uint32_t)(width / 16 - 1));
    bitwriter_write_ue(&bw, (uint32_t)(height / 16 - 1));
    bitwriter_write_bits(&bw, 0xC, 4);                /* frame_mbs_only=1, direct_8x8=1, no crop, no VUI */
    bitwriter_write_trailing_bits(&bw);
    return bitwriter_get_size(&bw);
}

/* PPS (src/h264_writer.c:105-127) */
size_t h264_generate_pps(uint8_t *rbsp, size_t capacity) {
    BitWriter bw;
    bitwriter_init(&bw, rbsp, capacity);
    bitwriter_write_ue(&bw, 0);                       /* pps id */
    bitwriter_write_ue(&bw, 0);                       /* sps id */
    bitwriter_write_bits(&bw, 0, 2);                  /* CAVLC, no bottom-field POC */
    bitwriter_write_ue(&bw, 0);                       /* one slice group */
    bitwriter_write_ue(&bw, 1);                       /* two default L0 refs */
    bitwriter_write_ue(&bw, 0);
    bitwriter_write_bits(&bw, 0, 3);                  /* weighted_pred_flag, weighted_bipred_idc */
    bitwriter_write_se(&bw, 0);
    bitwriter_write_se(&bw, 0);
    bitwriter_write_se(&bw, 0);
    bitwriter_write_bits(&bw, 0x4, 3);                /* deblocking control present, no constrained intra, no redundant pics */
    bitwriter_write_trailing_bits(&bw);
    return bitwriter_get_size(&bw);
}

/* ------------------------------------------------------------ I-frame rewrite */

typedef struct {
    size_t payload_bit;
    int32_t qp_delta;
    uint32_t deblock_idc;
    int32_t alpha, beta;
} IdrHeader;

/* IDR slice header of the atlas encoder (src/h264_writer.c:194-226) */
static void read_idr_header(const uint8_t *rbsp, size_t n, const ComposerConfig *cfg, IdrHeader *h) {
    BitReader br;
    bitreader_init(&br, rbsp, n);
    memset(h, 0, sizeof *h);
    bitreader_read_ue(&br);                           /* first_mb */
    bitreader_read_ue(&br);                           /* slice_type */
    bitreader_read_ue(&br);                           /* pps id */
    bitreader_read_bits(&br, cfg->log2_max_frame_num);
    bitreader_read_ue(&br);                           /* idr_pic_id */
    if (cfg->pic_order_cnt_type == 0) bitreader_read_bits(&br, cfg->log2_max_pic_order_cnt_lsb);
    bitreader_read_bits(&br, 2);                      /* no_output_of_prior_pics, long_term_reference */
    h->qp_delta = bitreader_read_se(&br);
    if (cfg->deblocking_filter_control_present_flag) {
        h->deblock_idc = bitreader_read_ue(&br);
        if (h->deblock_idc != 1) { h->alpha = bitreader_read_se(&br); h->beta = bitreader_read_se(&br); }
    }
    h->payload_bit = bitreader_get_bit_position(&br);
}

/* append src bits [start, start+n) (copy_bits, src/h264_writer.c:228-240), 32 bits per step */
static void append_bits(BitWriter *bw, const uint8_t *src, size_t src_size, size_t start, size_t n) {
    size_t p = start;
    while (n > 0) {
        int take = n >= 32 ? 32 : (int)n;
        uint64_t window = 0;
        size_t byte = p >> 3;
        for (int k = 0; k < 5; k++) window = (window << 8) | (byte + k < src_size ? src[byte + k] : 0);
        uint32_t v = (uint32_t)(window >> (8 - (p & 7)));        /* 32 bits starting at p */
        bitwriter_write_bits(bw, v >> (32 - take), take);
        p += (size_t)take;
        n -= (size_t)take;
    }
}

static void write_tail_of_i_header(BitWriter *bw, const ComposerConfig *cfg, const IdrHeader *h) {
    bitwriter_write_se(bw, h->qp_delta);
    if (cfg->deblocking_filter_control_present_flag) {
        bitwriter_write_ue(bw, h->deblock_idc);
        if (h->deblock_idc != 1) { bitwriter_write_se(bw, h->alpha); bitwriter_write_se(bw, h->beta); }
    }
}

/* RefA: IDR with long_term_reference_flag = 1 (src/h264_writer.c:242-294) */
size_t h264_rewrite_idr_frame(NALWriter *nw, ComposerConfig *write_cfg, ComposerConfig *parse_cfg,
                              const uint8_t *rbsp, size_t rbsp_size) {
    IdrHeader h;
    read_idr_header(rbsp, rbsp_size, parse_cfg, &h);
    size_t cap = rbsp_size + 256;
    uint8_t *tmp = malloc(cap);
    if (!tmp) return 0;
    BitWriter bw;
    bitwriter_init(&bw, tmp, cap);
    bitwriter_write_ue(&bw, 0);
    bitwriter_write_ue(&bw, SLICE_TYPE_I_ALL);
    bitwriter_write_ue(&bw, 0);
    bitwriter_write_bits(&bw, 0, write_cfg->log2_max_frame_num);
    bitwriter_write_ue(&bw, (uint32_t)write_cfg->idr_pic_id);
    if (write_cfg->pic_order_cnt_type == 0) bitwriter_write_bits(&bw, 0, write_cfg->log2_max_pic_order_cnt_lsb);
    bitwriter_write_bits(&bw, 1, 2);                  /* no_output_of_prior_pics = 0, long_term_reference = 1 */
    write_tail_of_i_header(&bw, write_cfg, &h);
    append_bits(&bw, rbsp, rbsp_size, h.payload_bit, rbsp_size * 8 - h.payload_bit);
    size_t w = nal_write_unit(nw, NAL_REF_IDC_HIGHEST, NAL_TYPE_IDR, tmp, bitwriter_get_size(&bw), 1);
    free(tmp);
    write_cfg->frame_num = 1;
    return w;
}

/* RefB: non-IDR I slice, MMCO 4 then 6 -> long-term index 1 (src/h264_writer.c:296-350) */
size_t h264_rewrite_as_non_idr_i_frame(NALWriter *nw, ComposerConfig *write_cfg, ComposerConfig *parse_cfg,
                                       const uint8_t *rbsp, size_t rbsp_size, int frame_num) {
    IdrHeader h;
    read_idr_header(rbsp, rbsp_size, parse_cfg, &h);
    size_t cap = rbsp_size + 256;
    uint8_t *tmp = malloc(cap);
    if (!tmp) return 0;
    BitWriter bw;
    bitwriter_init(&bw, tmp, cap);
    bitwriter_write_ue(&bw, 0);
    bitwriter_write_ue(&bw, SLICE_TYPE_I_ALL);
    bitwriter_write_ue(&bw, 0);
    bitwriter_write_bits(&bw, (uint32_t)frame_num, write_cfg->log2_max_frame_num);
    if (write_cfg->pic_order_cnt_type == 0)
        bitwriter_write_bits(&bw, (uint32_t)frame_num * 2, write_cfg->log2_max_pic_order_cnt_lsb);
    bitwriter_write_bit(&bw, 1);                      /* adaptive marking */
    bitwriter_write_ue(&bw, 4); bitwriter_write_ue(&bw, 2);
    bitwriter_write_ue(&bw, 6); bitwriter_write_ue(&bw, 1);
    bitwriter_write_ue(&bw, 0);
    write_tail_of_i_header(&bw, write_cfg, &h);
    append_bits(&bw, rbsp, rbsp_size, h.payload_bit, rbsp_size * 8 - h.payload_bit);
    size_t w = nal_write_unit(nw, NAL_REF_IDC_HIGHEST, NAL_TYPE_SLICE, tmp, bitwriter_get_size(&bw), 1);
    free(tmp);
    write_cfg->frame_num = frame_num + 1;
    return w;
}

/* ------------------------------------------------------------ P-slice headers */

/* src/h264_writer.c:455-539; which variant is used follows :549-553 and :687 */
int h264_p_slice_header_bits(const ComposerConfig *cfg, int waypoint_frame, uint64_t hdr[4]) {
    uint8_t buf[40];
    BitWriter bw;
    memset(buf, 0, sizeof buf);
    bitwriter_init(&bw, buf, sizeof buf);
    int fn_bits = cfg->log2_max_frame_num;
    int fn = cfg->frame_num % (1 << fn_bits);
    bitwriter_write_bits(&bw, 0x7, 3);                /* first_mb ue(0), slice_type P ue(0), pps ue(0) */
    bitwriter_write_bits(&bw, (uint32_t)fn, fn_bits);
    if (cfg->pic_order_cnt_type == 0) {
        int pb = cfg->log2_max_pic_order_cnt_lsb;
        bitwriter_write_bits(&bw, (uint32_t)(fn * 2) & ((1u << pb) - 1u), pb);
    }
    bitwriter_write_bit(&bw, 1);                      /* num_ref_idx_active_override_flag */
    int list_waypoints = waypoint_frame || cfg->num_waypoints > 0;
    bitwriter_write_ue(&bw, list_waypoints ? (uint32_t)(1 + cfg->num_waypoints) : 1u);
    bitwriter_write_bit(&bw, 1);                      /* ref_pic_list_modification_flag_l0 */
    for (uint32_t lt = 0; lt < 2; lt++) { bitwriter_write_ue(&bw, 2); bitwriter_write_ue(&bw, lt); }
    if (list_waypoints)
        for (int i = 0; i < cfg->num_waypoints; i++)
            if (cfg->waypoints[i].valid) {
                bitwriter_write_ue(&bw, 2);
                bitwriter_write_ue(&bw, (uint32_t)cfg->waypoints[i].long_term_idx);
            }
    bitwriter_write_ue(&bw, 3);
    if (waypoint_frame) {                             /* reference picture: mark it long-term */
        uint32_t lt = 2u + (uint32_t)cfg->num_waypoints;
        bitwriter_write_bit(&bw, 1);
        bitwriter_write_ue(&bw, 4); bitwriter_write_ue(&bw, lt + 1);
        bitwriter_write_ue(&bw, 6); bitwriter_write_ue(&bw, lt);
        bitwriter_write_ue(&bw, 0);
    }
    bitwriter_write_se(&bw, 0);                       /* slice_qp_delta */
    if (cfg->deblocking_filter_control_present_flag) bitwriter_write_ue(&bw, 1);
    int nbits = (int)bitwriter_get_bit_position(&bw);
    bitwriter_get_size(&bw);                          /* left-align the pending bits into buf */
    assert(nbits <= H264GPU_MAX_HDR_BITS);
    for (int w = 0; w < 4; w++) {
        uint64_t v = 0;
        for (int k = 0; k < 8; k++) v = (v << 8) | buf[8 * w + k];
        hdr[w] = v;
    }
    return nbits;
}

/* ------------------------------------------------------------ waypoints, regions */

int h264_needs_waypoint(ComposerConfig *cfg, int offset_px) {
    if (offset_px == 0 || offset_px % MV_LIMIT_PX) return 0;
    for (int i = 0; i < cfg->num_waypoints; i++)
        if (cfg->waypoints[i].valid && cfg->waypoints[i].offset_px == offset_px) return 0;
    return 1;
}

void h264_scroll_regions(const ComposerConfig *cfg, int offset_px, int waypoint_frame, int out[5]) {
    int top_ref = 0, top_mv = offset_px;
    int bot_ref = 1, bot_mv = offset_px - cfg->height;
    if (offset_px > MV_LIMIT_PX) {
        /* closest registered waypoint at or below the offset, within reach (:559-571, :693-705) */
        int best = 0;
        for (int i = 0; i < cfg->num_waypoints; i++) {
            const WaypointInfo *w = &cfg->waypoints[i];
            if (w->valid && w->offset_px <= offset_px && w->offset_px > best &&
                offset_px - w->offset_px <= MV_LIMIT_PX) {
                best = w->offset_px;
                top_ref = 2 + i;
                top_mv = offset_px - w->offset_px;
            }
        }
    }
    if (!waypoint_frame && bot_mv < -MV_LIMIT_PX) {
        /* first waypoint above the offset that B can reach (:575-588); never in a waypoint frame (:726-729) */
        for (int i = 0; i < cfg->num_waypoints; i++) {
            const WaypointInfo *w = &cfg->waypoints[i];
            if (w->valid && w->offset_px > offset_px && offset_px - w->offset_px >= -MV_LIMIT_PX) {
                bot_ref = 2 + i;
                bot_mv = offset_px - w->offset_px;
                break;
            }
        }
    }
    out[0] = (cfg->height - offset_px) / 16;          /* first row of the bottom band (:555) */
    out[1] = top_ref; out[2] = top_mv;
    out[3] = bot_ref; out[4] = bot_mv;
}

static void describe_frame(const ComposerConfig *cfg, int offset_px, int waypoint_frame, LegacyFrame *f) {
    int r[5];
    f->hdr_nbits = h264_p_slice_header_bits(cfg, waypoint_frame, f->hdr);
    f->nal_header = ((waypoint_frame ? NAL_REF_IDC_HIGH : NAL_REF_IDC_NONE) << 5) | NAL_TYPE_SLICE;
    f->num_refs = 2 + cfg->num_waypoints;             /* :633 */
    h264_scroll_regions(cfg, offset_px, waypoint_frame, r);
    f->split_row = r[0];
    f->ref_top = r[1]; f->mv_top_qpel = r[2] * 4;     /* :619-620 */
    f->ref_bot = r[3]; f->mv_bot_qpel = r[4] * 4;
}

static void register_waypoint(ComposerConfig *cfg, int offset_px) {
    if (cfg->num_waypoints < MAX_WAYPOINTS) {         /* saturates at eight (:771-777) */
        WaypointInfo *w = &cfg->waypoints[cfg->num_waypoints];
        w->offset_px = offset_px;
        w->long_term_idx = 2 + cfg->num_waypoints;
        w->valid = 1;
        cfg->num_waypoints++;
    }
}

/* ------------------------------------------------------------ P-frame writers */

static size_t write_one(NALWriter *nw, ComposerConfig *cfg, int offset_px, int waypoint_frame) {
    LegacyFrame f;
    uint32_t size = 0;
    size_t total = 0;
    describe_frame(cfg, offset_px, waypoint_frame, &f);
    int rc = gpu_session_compose_legacy(&f, 1, cfg->mb_width, cfg->mb_height,
                                        nw->output + nw->output_pos,
                                        nw->output_capacity - nw->output_pos, &size, &total);
    if (rc == H264GPU_ERR_OVERFLOW) {                 /* the reference asserts here (src/nal.c:34,59) */
        fprintf(stderr, "h264_writer: output buffer overflow\n");
        abort();
    }
    if (rc) return 0;
    nw->output_pos += total;
    if (waypoint_frame) register_waypoint(cfg, offset_px);
    cfg->frame_num++;                                 /* :662, :780 */
    return total;
}

size_t h264_write_scroll_p_frame(NALWriter *nw, ComposerConfig *cfg, int offset_px) {
    return write_one(nw, cfg, offset_px, 0);
}

size_t h264_write_waypoint_p_frame(NALWriter *nw, ComposerConfig *cfg, int offset_px) {
    return write_one(nw, cfg, offset_px, 1);
}

/*
 * The composer's per-frame sequence (src/composer.c:255-264) for many offsets
 * in one GPU batch: the host walks the stream state to build every frame's
 * header and regions, the device does the rest.  Returns frames written
 * (waypoint frames included) or -1.
 */
int h264_write_scroll_sequence(NALWriter *nw, ComposerConfig *cfg, const int *offsets, int n,
                               int *waypoints_made) {
    LegacyFrame *fr = malloc(sizeof(LegacyFrame) * 2 * (size_t)(n > 0 ? n : 1));
    uint32_t *sizes = malloc(sizeof(uint32_t) * 2 * (size_t)(n > 0 ? n : 1));
    if (!fr || !sizes) { free(fr); free(sizes); return -1; }
    ComposerConfig st = *cfg;                         /* evolve a copy; commit on success */
    int k = 0, made = 0;
    for (int i = 0; i < n; i++) {
        if (h264_needs_waypoint(&st, offsets[i])) {
            describe_frame(&st, offsets[i], 1, &fr[k++]);
            register_waypoint(&st, offsets[i]);
            st.frame_num++;
            made++;
        }
        describe_frame(&st, offsets[i], 0, &fr[k++]);
        st.frame_num++;
    }
    size_t total = 0;
    int rc = gpu_session_compose_legacy(fr, k, cfg->mb_width, cfg->mb_height,
                                        nw->output + nw->output_pos,
                                        nw->output_capacity - nw->output_pos, sizes, &total);
    free(fr);
    free(sizes);
    if (rc == H264GPU_ERR_OVERFLOW) {
        fprintf(stderr, "h264_writer: output buffer overflow\n");
        abort();
    }
    if (rc) return -1;
    nw->output_pos += total;
    *cfg = st;
    if (waypoints_made) *waypoints_made = made;
    return k;
}

/*
 * The same sequence with UI hints on top of every scroll frame (include/composer.h: ComposerFrameHints;
 * docs/MASTER_DESIGN.md:105-133,142-171).  Waypoint frames stay plain scroll frames.
 */
#include "composer.h"
int h264_write_hinted_sequence(void *nal_writer, void *composer_config, const void *frames_v, int n, int *waypoints_made) {
    NALWriter *nw = nal_writer;
    ComposerConfig *cfg = composer_config;
    const ComposerFrameHints *hf = frames_v;
    if (n <= 0) return 0;
    LegacyFrame *fr = malloc(sizeof(LegacyFrame) * 2 * (size_t)n);
    HintOverlay *ov = calloc(2 * (size_t)n, sizeof(HintOverlay));
    uint32_t *sizes = malloc(sizeof(uint32_t) * 2 * (size_t)n);
    if (!fr || !ov || !sizes) { free(fr); free(ov); free(sizes); return -1; }
    ComposerConfig st = *cfg;                         /* evolve a copy; commit on success */
    int k = 0, made = 0;
    for (int i = 0; i < n; i++) {
        if (h264_needs_waypoint(&st, hf[i].offset_px)) {
            describe_frame(&st, hf[i].offset_px, 1, &fr[k++]);
            register_waypoint(&st, hf[i].offset_px);
            st.frame_num++;
            made++;
        }
        describe_frame(&st, hf[i].offset_px, 0, &fr[k]);
        ov[k].regions = hf[i].regions;
        ov[k].n_regions = hf[i].regions ? hf[i].n_regions : 0;
        ov[k].dyn = hf[i].dyn;
        k++;
        st.frame_num++;
    }
    size_t total = 0;
    int rc = gpu_session_compose_hinted(fr, ov, k, cfg->mb_width, cfg->mb_height, nw->output + nw->output_pos,
                                        nw->output_capacity - nw->output_pos, sizes, &total);
    free(fr); free(ov); free(sizes);
    if (rc == H264GPU_ERR_OVERFLOW) {                 /* the reference asserts here (src/nal.c:34,59) */
        fprintf(stderr, "h264_writer: output buffer overflow\n");
        abort();
    }
    if (rc) return -1;
    nw->output_pos += total;
    *cfg = st;
    if (waypoints_made) *waypoints_made = made;
    return k;
}
