/*
 * nal.c - host Annex-B framing behind include/nal.h (reference: src/nal.c).
 * Used for the once-per-stream NAL units; P-slices are framed on the GPU.
 */
#include "nal.h"
#include <assert.h>
#include <string.h>

void nal_writer_init(NALWriter *nw, uint8_t *output, size_t output_capacity,
                     uint8_t *rbsp_temp, size_t rbsp_capacity) {
    nw->output = output;
    nw->output_capacity = output_capacity;
    nw->output_pos = 0;
    nw->rbsp = rbsp_temp;
    nw->rbsp_capacity = rbsp_capacity;
}

/*
 * Emulation prevention (src/nal.c:24-50): a 0x03 goes in front of any byte <= 3
 * that follows two zero bytes; the zero count restarts after an insertion.
 * Runs without zero bytes are block-copied.
 */
size_t rbsp_to_ebsp(uint8_t *ebsp, size_t ebsp_capacity, const uint8_t *rbsp, size_t rbsp_size) {
    size_t o = 0, i = 0;
    int zeros = 0;
    while (i < rbsp_size) {
        if (zeros == 0) {
            const uint8_t *z = memchr(rbsp + i, 0, rbsp_size - i);
            size_t span = z ? (size_t)(z - (rbsp + i)) : rbsp_size - i;
            if (span) {
                assert(o + span <= ebsp_capacity);      /* src/nal.c:39 */
                memcpy(ebsp + o, rbsp + i, span);
                o += span; i += span;
                continue;
            }
        }
        uint8_t v = rbsp[i++];
        if (zeros >= 2 && v <= 3) {
            assert(o < ebsp_capacity);                  /* src/nal.c:34 */
            ebsp[o++] = 3;
            zeros = 0;
        }
        assert(o < ebsp_capacity);
        ebsp[o++] = v;
        zeros = v == 0 ? zeros + 1 : 0;
    }
    return o;
}

size_t nal_write_unit(NALWriter *nw, int nal_ref_idc, int nal_type,
                      const uint8_t *rbsp, size_t rbsp_size, int use_long_startcode) {
    size_t begin = nw->output_pos;
    size_t sc = use_long_startcode ? 4 : 3;
    assert(nw->output_pos + sc + 1 <= nw->output_capacity);   /* src/nal.c:59,65,73 */
    uint8_t *p = nw->output + nw->output_pos;
    memset(p, 0, sc - 1);
    p[sc - 1] = 1;
    p[sc] = (uint8_t)(((nal_ref_idc & 3) << 5) | (nal_type & 31));
    nw->output_pos += sc + 1;
    nw->output_pos += rbsp_to_ebsp(nw->output + nw->output_pos,
                                   nw->output_capacity - nw->output_pos, rbsp, rbsp_size);
    return nw->output_pos - begin;
}

size_t nal_writer_get_size(NALWriter *nw) { return nw->output_pos; }
uint8_t *nal_writer_get_output(NALWriter *nw) { return nw->output; }
