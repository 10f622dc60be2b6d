/*
 * nal_parser.c - Annex-B scanner + partial SPS/PPS parser behind
 * include/nal_parser.h (reference: src/nal_parser.c).  Init-time only.
 */
#include "nal_parser.h"
#include "bitwriter.h"

void nal_parser_init(NALParser *parser, const uint8_t *data, size_t size) {
    parser->data = data;
    parser->size = size;
    parser->pos = 0;
}

/* length of the start code beginning at i (3 or 4), or 0 (src/nal_parser.c:14-26) */
static size_t start_code_at(const uint8_t *d, size_t size, size_t i) {
    if (i + 2 < size && d[i] == 0 && d[i + 1] == 0) {
        if (d[i + 2] == 1) return 3;
        if (i + 3 < size && d[i + 2] == 0 && d[i + 3] == 1) return 4;
    }
    return 0;
}

int nal_parser_next(NALParser *parser, NALUnit *unit) {
    const uint8_t *d = parser->data;
    size_t size = parser->size, i = parser->pos, begin = size;
    for (; i + 2 < size; i++) {
        size_t sc = start_code_at(d, size, i);
        if (sc) { begin = i + sc; break; }
    }
    if (begin >= size) return 0;
    size_t end = size;
    for (i = begin; i + 2 < size; i++)
        if (start_code_at(d, size, i)) { end = i; break; }
    while (end > begin && d[end - 1] == 0) end--;        /* trailing zeros belong to the next start code */
    if (end <= begin) { parser->pos = end; return 0; }
    unit->nal_ref_idc = (d[begin] >> 5) & 3;
    unit->nal_unit_type = d[begin] & 31;
    unit->data = d + begin + 1;
    unit->size = end - begin - 1;
    parser->pos = end;
    return 1;
}

/* drop 0x03 after two zeros when the next byte is <= 3 (src/nal_parser.c:67-88) */
size_t ebsp_to_rbsp(uint8_t *rbsp, const uint8_t *ebsp, size_t ebsp_size) {
    size_t o = 0;
    int zeros = 0;
    for (size_t i = 0; i < ebsp_size; i++) {
        uint8_t v = ebsp[i];
        if (zeros >= 2 && v == 3 && i + 1 < ebsp_size && ebsp[i + 1] <= 3) { zeros = 0; continue; }
        rbsp[o++] = v;
        zeros = v == 0 ? zeros + 1 : 0;
    }
    return o;
}

static int is_high_profile(int p) {
    static const int hp[] = {100, 110, 122, 244, 44, 83, 86, 118, 128, 138, 139, 134};
    for (unsigned k = 0; k < sizeof hp / sizeof hp[0]; k++) if (hp[k] == p) return 1;
    return 0;
}

/* src/nal_parser.c:137-222 */
int parse_sps(const uint8_t *rbsp, size_t size, int *width, int *height,
              int *log2_max_frame_num, int *pic_order_cnt_type, int *log2_max_pic_order_cnt_lsb) {
    BitReader br;
    bitreader_init(&br, rbsp, size);
    int profile = (int)bitreader_read_bits(&br, 8);
    bitreader_read_bits(&br, 16);                        /* constraint flags, level */
    bitreader_read_ue(&br);                              /* sps id */
    if (is_high_profile(profile)) {
        if (bitreader_read_ue(&br) == 3) bitreader_read_bit(&br);
        bitreader_read_ue(&br);
        bitreader_read_ue(&br);
        bitreader_read_bit(&br);
        if (bitreader_read_bit(&br)) return -1;          /* scaling lists unsupported */
    }
    *log2_max_frame_num = (int)bitreader_read_ue(&br) + 4;
    *pic_order_cnt_type = (int)bitreader_read_ue(&br);
    *log2_max_pic_order_cnt_lsb = 0;
    if (*pic_order_cnt_type == 0) *log2_max_pic_order_cnt_lsb = (int)bitreader_read_ue(&br) + 4;
    else if (*pic_order_cnt_type == 1) { bitreader_read_bit(&br); return -1; }
    bitreader_read_ue(&br);                              /* max_num_ref_frames */
    bitreader_read_bit(&br);                             /* gaps allowed */
    int w_mbs = (int)bitreader_read_ue(&br) + 1;
    int h_units = (int)bitreader_read_ue(&br) + 1;
    if (!bitreader_read_bit(&br)) { bitreader_read_bit(&br); h_units *= 2; }
    *width = w_mbs * 16;
    *height = h_units * 16;
    return 0;
}

/* src/nal_parser.c:224-276 */
int parse_pps(const uint8_t *rbsp, size_t size, int *num_ref_idx_l0_default_minus1,
              int *deblocking_filter_control_present_flag) {
    BitReader br;
    bitreader_init(&br, rbsp, size);
    bitreader_read_ue(&br);
    bitreader_read_ue(&br);
    bitreader_read_bits(&br, 2);                         /* entropy mode, bottom_field_pic_order */
    if (bitreader_read_ue(&br) > 0) return -1;           /* slice groups unsupported */
    *num_ref_idx_l0_default_minus1 = (int)bitreader_read_ue(&br);
    bitreader_read_ue(&br);
    bitreader_read_bits(&br, 3);                         /* weighted_pred_flag, weighted_bipred_idc */
    bitreader_read_ue(&br);                              /* three se() fields, values unused */
    bitreader_read_ue(&br);
    bitreader_read_ue(&br);
    *deblocking_filter_control_present_flag = bitreader_read_bit(&br);
    return 0;
}
