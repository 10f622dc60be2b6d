/*
 * bitwriter.c - host bit writer/reader behind include/bitwriter.h.
 *
 * Same observable behaviour as the reference's src/bitwriter.c (struct state,
 * Exp-Golomb codes, trailing bits, overflow = assert) but whole fields are
 * merged into the pending byte at once instead of one call per bit.
 */
#include "bitwriter.h"
#include <assert.h>

void bitwriter_init(BitWriter *bw, uint8_t *buffer, size_t capacity) {
    bw->buffer = buffer;
    bw->capacity = capacity;
    bw->byte_pos = 0;
    bw->bit_pos = 0;
    bw->current_byte = 0;
}

/* append the low n bits of value (n in 0..32), most significant first */
static void put_bits(BitWriter *bw, uint32_t value, int n) {
    while (n > 0) {
        int room = 8 - bw->bit_pos;
        int take = n < room ? n : room;
        uint32_t piece = (value >> (n - take)) & ((1u << take) - 1u);
        bw->current_byte = (uint8_t)((bw->current_byte << take) | piece);
        bw->bit_pos += take;
        n -= take;
        if (bw->bit_pos == 8) {
            assert(bw->byte_pos < bw->capacity);       /* src/bitwriter.c:18 */
            bw->buffer[bw->byte_pos++] = bw->current_byte;
            bw->current_byte = 0;
            bw->bit_pos = 0;
        }
    }
}

void bitwriter_write_bit(BitWriter *bw, int bit) { put_bits(bw, (uint32_t)bit & 1u, 1); }

void bitwriter_write_bits(BitWriter *bw, uint32_t value, int n) {
    assert(n >= 1 && n <= 32);                          /* src/bitwriter.c:26 */
    put_bits(bw, value, n);
}

/* ue(v): floor(log2(v+1)) zeros, then v+1 (src/bitwriter.c:50-74) */
void bitwriter_write_ue(BitWriter *bw, uint32_t value) {
    uint32_t v = value + 1u;
    int m = 31 - __builtin_clz(v | 1u);
    if (v == 0) { put_bits(bw, 0, 1); return; }         /* value 0xFFFFFFFF wraps, as in the reference */
    put_bits(bw, 0, m);
    put_bits(bw, v, m + 1);
}

/* se(v): v > 0 -> 2v-1, otherwise -2v (src/bitwriter.c:91-101) */
void bitwriter_write_se(BitWriter *bw, int32_t value) {
    uint32_t k = value > 0 ? (uint32_t)value * 2u - 1u : (uint32_t)(-(int64_t)value) * 2u;
    bitwriter_write_ue(bw, k);
}

void bitwriter_write_trailing_bits(BitWriter *bw) {
    put_bits(bw, 1, 1);
    if (bw->bit_pos) put_bits(bw, 0, 8 - bw->bit_pos);
}

void bitwriter_flush(BitWriter *bw) {
    if (bw->bit_pos) put_bits(bw, 0, 8 - bw->bit_pos);
}

/* size in bytes; a partial byte is stored left-aligned but stays pending (src/bitwriter.c:124-131) */
size_t bitwriter_get_size(BitWriter *bw) {
    if (bw->bit_pos > 0) {
        bw->buffer[bw->byte_pos] = (uint8_t)(bw->current_byte << (8 - bw->bit_pos));
        return bw->byte_pos + 1;
    }
    return bw->byte_pos;
}

size_t bitwriter_get_bit_position(BitWriter *bw) { return bw->byte_pos * 8 + (size_t)bw->bit_pos; }
int bitwriter_is_byte_aligned(BitWriter *bw) { return bw->bit_pos == 0; }

/* ---------------------------------------------------------------- reader */

void bitreader_init(BitReader *br, const uint8_t *buffer, size_t size) {
    br->buffer = buffer;
    br->size = size;
    br->byte_pos = 0;
    br->bit_pos = 0;
}

int bitreader_read_bit(BitReader *br) {
    if (br->byte_pos >= br->size) return 0;             /* reads past the end give 0 */
    int bit = (br->buffer[br->byte_pos] >> (7 - br->bit_pos)) & 1;
    if (++br->bit_pos == 8) { br->bit_pos = 0; br->byte_pos++; }
    return bit;
}

uint32_t bitreader_read_bits(BitReader *br, int n) {
    uint32_t v = 0;
    while (n-- > 0) v = (v << 1) | (uint32_t)bitreader_read_bit(br);
    return v;
}

uint32_t bitreader_read_ue(BitReader *br) {
    int zeros = 0;
    while (bitreader_read_bit(br) == 0 && zeros < 32) zeros++;
    if (zeros == 0) return 0;
    return (1u << zeros) - 1u + bitreader_read_bits(br, zeros);
}

int32_t bitreader_read_se(BitReader *br) {
    uint32_t k = bitreader_read_ue(br);
    return (k & 1u) ? (int32_t)((k + 1u) / 2u) : -(int32_t)(k / 2u);
}

size_t bitreader_get_bit_position(BitReader *br) { return br->byte_pos * 8 + (size_t)br->bit_pos; }
int bitreader_is_byte_aligned(BitReader *br) { return br->bit_pos == 0; }

size_t bitreader_get_remaining_bytes(BitReader *br) {
    return br->size - br->byte_pos - (br->bit_pos != 0 ? 1 : 0);
}

const uint8_t *bitreader_get_pointer(BitReader *br) { return br->buffer + br->byte_pos; }
