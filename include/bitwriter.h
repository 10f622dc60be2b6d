/*
 * bitwriter.h - MSB-first bit writer/reader kept for the host-side syntax
 * (SPS, PPS, slice headers).  ABI-compatible with the reference's
 * include/bitwriter.h:17-91: same struct layouts, same 10 + 9 entry points,
 * so existing callers compile and link unchanged.  The per-macroblock bit
 * emission of P-frames no longer goes through this writer: it is computed on
 * the GPU (include/composer_gpu.h).
 */
#ifndef BITWRITER_H
#define BITWRITER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint8_t *buffer;
    size_t capacity;
    size_t byte_pos;
    int bit_pos;              /* bits already used in current_byte, 0..7 */
    uint8_t current_byte;
} BitWriter;

void   bitwriter_init(BitWriter *bw, uint8_t *buffer, size_t capacity);
void   bitwriter_write_bits(BitWriter *bw, uint32_t value, int n);   /* n in 1..32 */
void   bitwriter_write_bit(BitWriter *bw, int bit);
void   bitwriter_write_ue(BitWriter *bw, uint32_t value);
void   bitwriter_write_se(BitWriter *bw, int32_t value);
void   bitwriter_write_trailing_bits(BitWriter *bw);                 /* '1' then zeros to a byte edge */
void   bitwriter_flush(BitWriter *bw);
size_t bitwriter_get_size(BitWriter *bw);                            /* bytes, partial byte included */
size_t bitwriter_get_bit_position(BitWriter *bw);
int    bitwriter_is_byte_aligned(BitWriter *bw);

typedef struct {
    const uint8_t *buffer;
    size_t size;
    size_t byte_pos;
    int bit_pos;              /* bits consumed in the current byte, 0..7 */
} BitReader;

void     bitreader_init(BitReader *br, const uint8_t *buffer, size_t size);
uint32_t bitreader_read_bits(BitReader *br, int n);
int      bitreader_read_bit(BitReader *br);                          /* past the end: 0 */
uint32_t bitreader_read_ue(BitReader *br);
int32_t  bitreader_read_se(BitReader *br);
size_t   bitreader_get_bit_position(BitReader *br);
int      bitreader_is_byte_aligned(BitReader *br);
size_t   bitreader_get_remaining_bytes(BitReader *br);
const uint8_t *bitreader_get_pointer(BitReader *br);

#ifdef __cplusplus
}
#endif
#endif /* BITWRITER_H */
