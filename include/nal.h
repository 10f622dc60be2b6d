/*
 * nal.h - Annex-B NAL framing, ABI-compatible with the reference's
 * include/nal.h:10-73.  nal_write_unit / rbsp_to_ebsp remain host functions for
 * the once-per-stream units (SPS, PPS, rewritten I-frames); P-slices are framed
 * and emulation-protected inside the GPU kernel.
 */
#ifndef NAL_H
#define NAL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* nal_unit_type (H.264 Table 7-1) */
#define NAL_TYPE_SLICE 1
#define NAL_TYPE_IDR   5
#define NAL_TYPE_SEI   6
#define NAL_TYPE_SPS   7
#define NAL_TYPE_PPS   8
#define NAL_TYPE_AUD   9

/* nal_ref_idc */
#define NAL_REF_IDC_NONE    0
#define NAL_REF_IDC_LOW     1
#define NAL_REF_IDC_HIGH    2
#define NAL_REF_IDC_HIGHEST 3

typedef struct {
    uint8_t *output;          /* caller-owned Annex-B stream buffer */
    size_t output_capacity;
    size_t output_pos;
    uint8_t *rbsp;            /* caller-owned scratch */
    size_t rbsp_capacity;
} NALWriter;

void     nal_writer_init(NALWriter *nw, uint8_t *output, size_t output_capacity,
                         uint8_t *rbsp_temp, size_t rbsp_capacity);
/* start code (4 bytes when use_long_startcode, else 3) + header byte + EBSP; returns bytes appended */
size_t   nal_write_unit(NALWriter *nw, int nal_ref_idc, int nal_type,
                        const uint8_t *rbsp, size_t rbsp_size, int use_long_startcode);
size_t   nal_writer_get_size(NALWriter *nw);
uint8_t *nal_writer_get_output(NALWriter *nw);
/* insert 0x03 wherever 00 00 0x (x <= 3) would appear; returns the EBSP size */
size_t   rbsp_to_ebsp(uint8_t *ebsp, size_t ebsp_capacity, const uint8_t *rbsp, size_t rbsp_size);

#ifdef __cplusplus
}
#endif
#endif /* NAL_H */
