/*
 * nal_parser.h - Annex-B scanner and the partial SPS/PPS parser used when the
 * two reference atlases are loaded; ABI-compatible with the reference's
 * include/nal_parser.h:13-54.  Init-time only.
 */
#ifndef NAL_PARSER_H
#define NAL_PARSER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int nal_ref_idc;
    int nal_unit_type;
    const uint8_t *data;      /* payload after the header byte (EBSP) */
    size_t size;
    size_t rbsp_size;
} NALUnit;

typedef struct {
    const uint8_t *data;
    size_t size;
    size_t pos;
} NALParser;

void   nal_parser_init(NALParser *parser, const uint8_t *data, size_t size);
int    nal_parser_next(NALParser *parser, NALUnit *unit);            /* 1 = unit found, 0 = end */
size_t ebsp_to_rbsp(uint8_t *rbsp, const uint8_t *ebsp, size_t ebsp_size);
int    parse_sps(const uint8_t *rbsp, size_t size, int *width, int *height,
                 int *log2_max_frame_num, int *pic_order_cnt_type,
                 int *log2_max_pic_order_cnt_lsb);                   /* 0 ok, -1 unsupported */
int    parse_pps(const uint8_t *rbsp, size_t size, int *num_ref_idx_l0_default_minus1,
                 int *deblocking_filter_control_present_flag);

#ifdef __cplusplus
}
#endif
#endif /* NAL_PARSER_H */
