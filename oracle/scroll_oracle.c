/*
 * scroll_oracle.c - TEST INFRASTRUCTURE (see scroll_oracle.h).
 *
 * Plain-C restatement of the reference composer's P-slice path.  It is written
 * independently of the reference sources (word-at-a-time bit sink, closed-form
 * Exp-Golomb, hint-driven MB loop) but follows the reference algorithm, cited
 * per function as /root/reference file:line.
 */
#define _POSIX_C_SOURCE 199309L
#include "scroll_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ bits */

/* MSB-first bit sink (behaviour of src/bitwriter.c:13-32,124-131). */
typedef struct {
    uint8_t *buf;
    size_t cap;        /* bytes */
    size_t nbits;      /* bits appended so far */
    int overflow;
} Bits;

static void bits_init(Bits *b, uint8_t *buf, size_t cap) {
    b->buf = buf; b->cap = cap; b->nbits = 0; b->overflow = 0;
    if (cap) buf[0] = 0;
}

/* append the low n (0..32) bits of v, most significant first */
static void bits_put(Bits *b, uint32_t v, int n) {
    while (n > 0) {
        size_t byte = b->nbits >> 3;
        int used = (int)(b->nbits & 7);
        int room = 8 - used;
        int take = n < room ? n : room;
        if (byte >= b->cap) { b->overflow = 1; return; }
        if (used == 0) b->buf[byte] = 0;
        uint32_t chunk = (v >> (n - take)) & ((1u << take) - 1u);
        b->buf[byte] |= (uint8_t)(chunk << (room - take));
        b->nbits += (size_t)take;
        n -= take;
    }
}

static int ilog2_u32(uint32_t v) {           /* floor(log2 v), v >= 1 */
    int m = 0;
    while (v >>= 1) m++;
    return m;
}

/* ue(v): m zeros then (v+1) in m+1 bits, m = floor(log2(v+1))  (src/bitwriter.c:50-74) */
static int ue_len(uint32_t v) { return 2 * ilog2_u32(v + 1) + 1; }
static void bits_ue(Bits *b, uint32_t v) {
    int m = ilog2_u32(v + 1);
    if (m) bits_put(b, 0, m);
    bits_put(b, v + 1, m + 1);
}

/* se(v): v>0 -> 2v-1, else -2v, then ue  (src/bitwriter.c:91-101) */
static uint32_t se_map(int32_t v) { return v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v); }
static void bits_se(Bits *b, int32_t v) { bits_ue(b, se_map(v)); }

/* rbsp_stop_one_bit + zero alignment (src/bitwriter.c:103-111); returns bytes */
static size_t bits_finish(Bits *b) {
    bits_put(b, 1, 1);
    if (b->nbits & 7) bits_put(b, 0, 8 - (int)(b->nbits & 7));
    return b->nbits >> 3;
}

/* copy `n` bits of src starting at bit `start` (semantics of copy_bits,
 * src/h264_writer.c:228-240; reads past the end yield 0 like sbreader :156-157) */
static void bits_copy(Bits *b, const uint8_t *src, size_t src_bytes, size_t start, size_t n) {
    for (size_t i = 0; i < n; i++) {
        size_t p = start + i;
        uint32_t bit = (p >> 3) < src_bytes ? (src[p >> 3] >> (7 - (p & 7))) & 1u : 0u;
        bits_put(b, bit, 1);
    }
}

/* ------------------------------------------------------------ EBSP / NAL */

/* src/nal.c:24-50 */
size_t so_rbsp_to_ebsp(uint8_t *ebsp, size_t cap, const uint8_t *rbsp, size_t n) {
    size_t o = 0;
    int zeros = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t v = rbsp[i];
        if (zeros >= 2 && v <= 3) {
            if (o >= cap) return 0;
            ebsp[o++] = 3;
            zeros = 0;
        }
        if (o >= cap) return 0;
        ebsp[o++] = v;
        zeros = (v == 0) ? zeros + 1 : 0;
    }
    return o;
}

/* 4-byte start code, header byte, EBSP (src/nal.c:52-84 with use_long_startcode=1) */
static size_t nal_emit(uint8_t *out, size_t cap, int nal_header, const uint8_t *rbsp, size_t n) {
    if (cap < 5) return 0;
    out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 1;
    out[4] = (uint8_t)nal_header;
    size_t e = so_rbsp_to_ebsp(out + 5, cap - 5, rbsp, n);
    if (n && !e) return 0;
    return 5 + e;
}

/* ---------------------------------------------------------- MV predictor */

typedef struct { int mvx, mvy, ref, avail; } Nb;

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* What the reference's median3 really computes (src/h264_writer.c:362-367). */
static int ref_median(int a, int b, int c) { return imin(imax(a, b), c); }
static int true_median(int a, int b, int c) { return imax(imin(a, b), imin(imax(a, b), c)); }

static Nb nb_of(const H264GpuMbHint *h) {
    Nb n;
    n.avail = 1;
    if (h->kind == H264GPU_MB_DYN_INTRA) { n.mvx = 0; n.mvy = 0; n.ref = 255; }
    else { n.mvx = h->mv_x_qpel; n.mvy = h->mv_y_qpel; n.ref = h->ref_idx; }
    return n;
}

/* neighbours A (left), B (above), C (above-right, else D above-left)
 * (src/h264_writer.c:376-399) */
static void neighbours(const H264GpuMbHint *hints, int mb_w, int x, int y, Nb *a, Nb *b, Nb *c) {
    Nb none = {0, 0, 0, 0};
    *a = *b = *c = none;
    const H264GpuMbHint *row = hints + (size_t)y * mb_w;
    if (x > 0) *a = nb_of(&row[x - 1]);
    if (y > 0) {
        const H264GpuMbHint *up = row - mb_w;
        *b = nb_of(&up[x]);
        if (x + 1 < mb_w) *c = nb_of(&up[x + 1]);
        else if (x > 0)   *c = nb_of(&up[x - 1]);
    }
}

static void predict(int spec, Nb a, Nb b, Nb c, int ref, int *px, int *py) {
    if (spec) {
        /* H.264 8.4.1.3.1/.2: B and C unavailable, A available -> B = C = A */
        if (!b.avail && !c.avail && a.avail) { b = a; c = a; }
        int ma = a.avail && a.ref == ref, mb = b.avail && b.ref == ref, mc = c.avail && c.ref == ref;
        if (ma + mb + mc == 1) {
            const Nb *s = ma ? &a : (mb ? &b : &c);
            *px = s->mvx; *py = s->mvy;
        } else {
            *px = true_median(a.avail ? a.mvx : 0, b.avail ? b.mvx : 0, c.avail ? c.mvx : 0);
            *py = true_median(a.avail ? a.mvy : 0, b.avail ? b.mvy : 0, c.avail ? c.mvy : 0);
        }
        return;
    }
    /* reference rules in order (src/h264_writer.c:401-431) */
    int ma = a.avail && a.ref == ref, mb = b.avail && b.ref == ref, mc = c.avail && c.ref == ref;
    int n_av = a.avail + b.avail + c.avail;
    if (n_av == 0) { *px = 0; *py = 0; return; }
    if (n_av == 1) {
        const Nb *s = a.avail ? &a : (b.avail ? &b : &c);
        int m = a.avail ? ma : (b.avail ? mb : mc);
        *px = m ? s->mvx : 0; *py = m ? s->mvy : 0;
        return;
    }
    if (ma + mb + mc == 1) {
        const Nb *s = ma ? &a : (mb ? &b : &c);
        *px = s->mvx; *py = s->mvy;
        return;
    }
    *px = ref_median(a.avail ? a.mvx : 0, b.avail ? b.mvx : 0, c.avail ? c.mvx : 0);
    *py = ref_median(a.avail ? a.mvy : 0, b.avail ? b.mvy : 0, c.avail ? c.mvy : 0);
}

/* H.264 8.4.1.1 inferred P_Skip motion vector */
static void skip_mv(int spec, Nb a, Nb b, Nb c, int *sx, int *sy) {
    if (!a.avail || !b.avail ||
        (a.ref == 0 && a.mvx == 0 && a.mvy == 0) ||
        (b.ref == 0 && b.mvx == 0 && b.mvy == 0)) { *sx = 0; *sy = 0; return; }
    predict(spec, a, b, c, 0, sx, sy);
}

/* ------------------------------------------------------------- MB syntax */

/* ref_idx as te(): nothing / inverted bit / ue  (src/h264_writer.c:438-445) */
static int ref_len(int ref, int num_refs) {
    if (num_refs == 1) return 0;
    if (num_refs == 2) return 1;
    return ue_len((uint32_t)ref);
}
static void put_ref(Bits *b, int ref, int num_refs) {
    if (num_refs == 1) return;
    if (num_refs == 2) bits_put(b, (uint32_t)(1 - (ref & 1)), 1);
    else bits_ue(b, (uint32_t)ref);
}

/* mb_type ue(0) | ref | se(mvdx) | se(mvdy) | cbp ue(0)  (src/h264_writer.c:434-453) */
static void put_p16x16(Bits *b, int ref, int dx, int dy, int num_refs) {
    bits_put(b, 1, 1);
    put_ref(b, ref, num_refs);
    bits_se(b, dx);
    bits_se(b, dy);
    bits_put(b, 1, 1);
}
static int len_p16x16(int ref, int dx, int dy, int num_refs) {
    return 2 + ref_len(ref, num_refs) + ue_len(se_map(dx)) + ue_len(se_map(dy));
}

/* ---------------------------------------------------------- frame driver */

typedef struct {
    int coded;          /* 0 = P_Skip */
    int px, py;         /* predictor used for mvd */
} MbDecision;

static MbDecision decide(const H264GpuMbHint *hints, int mb_w, int x, int y, int flags) {
    const H264GpuMbHint *h = &hints[(size_t)y * mb_w + x];
    MbDecision d = {1, 0, 0};
    if (h->kind != H264GPU_MB_INTER) return d;
    int spec = (flags & H264GPU_PRED_SPEC) != 0;
    Nb a, b, c;
    neighbours(hints, mb_w, x, y, &a, &b, &c);
    predict(spec, a, b, c, h->ref_idx, &d.px, &d.py);
    if ((flags & H264GPU_SKIP_SPEC) && h->ref_idx == 0) {
        int sx, sy;
        skip_mv(spec, a, b, c, &sx, &sy);
        if (h->mv_x_qpel == sx && h->mv_y_qpel == sy) d.coded = 0;
    }
    return d;
}

/* outer loop of src/h264_writer.c:595-656 driven by hints */
static size_t frame_rbsp(const H264GpuMbHint *hints, int mb_w, int mb_h, int num_refs, int flags,
                         const uint64_t *hdr_bits, int hdr_nbits,
                         const uint8_t *payload, size_t payload_bytes,
                         const uint32_t *dyn_bit_off, uint8_t *rbsp, size_t cap) {
    Bits b;
    bits_init(&b, rbsp, cap);
    for (int k = 0; k < hdr_nbits; k++)
        bits_put(&b, (uint32_t)((hdr_bits[k >> 6] >> (63 - (k & 63))) & 1u), 1);

    uint32_t run = 0;
    for (int y = 0; y < mb_h; y++) {
        for (int x = 0; x < mb_w; x++) {
            const H264GpuMbHint *h = &hints[(size_t)y * mb_w + x];
            MbDecision d = decide(hints, mb_w, x, y, flags);
            if (!d.coded) { run++; continue; }
            bits_ue(&b, run);                       /* mb_skip_run (:630) */
            run = 0;
            if (h->kind == H264GPU_MB_INTER) {
                put_p16x16(&b, h->ref_idx, h->mv_x_qpel - d.px, h->mv_y_qpel - d.py, num_refs);
            } else {
                uint32_t b0 = dyn_bit_off[h->dyn_idx], b1 = dyn_bit_off[h->dyn_idx + 1];
                bits_copy(&b, payload, payload_bytes, b0, b1 - b0);
            }
        }
    }
    if (run) bits_ue(&b, run);                      /* trailing run (:648-650) */
    size_t n = bits_finish(&b);
    return b.overflow ? 0 : n;
}

size_t so_compose_frame(const H264GpuMbHint *hints, int mb_w, int mb_h,
                        int num_refs, int flags,
                        const uint64_t *hdr_bits, int hdr_nbits, int nal_header,
                        const uint8_t *payload, size_t payload_bytes,
                        const uint32_t *dyn_bit_off,
                        uint8_t *out, size_t out_cap) {
    /* worst case 104 bits per inter MB; spliced MBs bring their own length */
    size_t cap = 64 + (size_t)mb_w * mb_h * 16 + payload_bytes + 16;
    uint8_t *rbsp = malloc(cap);
    if (!rbsp) return 0;
    size_t n = frame_rbsp(hints, mb_w, mb_h, num_refs, flags, hdr_bits, hdr_nbits,
                          payload, payload_bytes, dyn_bit_off, rbsp, cap);
    size_t w = n ? nal_emit(out, out_cap, nal_header, rbsp, n) : 0;
    free(rbsp);
    return w;
}

void so_frame_stages(const H264GpuMbHint *hints, int mb_w, int mb_h,
                     int num_refs, int flags, const uint32_t *dyn_bit_off,
                     int16_t *pred_x, int16_t *pred_y, uint32_t *bitlen) {
    uint32_t run = 0;
    for (int y = 0; y < mb_h; y++)
        for (int x = 0; x < mb_w; x++) {
            size_t i = (size_t)y * mb_w + x;
            const H264GpuMbHint *h = &hints[i];
            MbDecision d = decide(hints, mb_w, x, y, flags);
            pred_x[i] = (int16_t)d.px; pred_y[i] = (int16_t)d.py;
            if (!d.coded) { run++; bitlen[i] = 0; continue; }
            uint32_t len = (uint32_t)ue_len(run);
            run = 0;
            if (h->kind == H264GPU_MB_INTER)
                len += (uint32_t)len_p16x16(h->ref_idx, h->mv_x_qpel - d.px,
                                            h->mv_y_qpel - d.py, num_refs);
            else
                len += dyn_bit_off[h->dyn_idx + 1] - dyn_bit_off[h->dyn_idx];
            bitlen[i] = len;
        }
}

/* ------------------------------------------------- host state and headers */

void so_stream_init(SoStream *st, int width, int height, int deblock_present) {
    memset(st, 0, sizeof(*st));
    st->width = width; st->height = height;
    st->mb_w = width / 16; st->mb_h = height / 16;       /* :17-18 */
    st->log2_max_frame_num = 4; st->poc_type = 2; st->log2_max_poc_lsb = 4;  /* src/composer.c:198-201 */
    st->deblock_present = deblock_present;
    st->frame_num = 2;          /* after the IDR and the non-IDR I frame (:292,:348) */
}

int so_slice_header(const SoStream *st, int waypoint_frame, uint64_t hdr[4]) {
    uint8_t tmp[32];
    Bits b;
    memset(tmp, 0, sizeof tmp);
    bits_init(&b, tmp, sizeof tmp);
    int fn = st->frame_num % (1 << st->log2_max_frame_num);        /* :546-547 */
    bits_ue(&b, 0);                       /* first_mb_in_slice */
    bits_ue(&b, 0);                       /* slice_type P */
    bits_ue(&b, 0);                       /* pps id */
    bits_put(&b, (uint32_t)fn, st->log2_max_frame_num);
    if (st->poc_type == 0)
        bits_put(&b, (uint32_t)(fn * 2) & ((1u << st->log2_max_poc_lsb) - 1u), st->log2_max_poc_lsb);
    bits_put(&b, 1, 1);                   /* num_ref_idx_active_override_flag */
    int with_wp = waypoint_frame || st->n_wp > 0;                   /* :549-553 */
    bits_ue(&b, with_wp ? (uint32_t)(2 + st->n_wp - 1) : 1u);
    bits_put(&b, 1, 1);                   /* ref_pic_list_modification_flag_l0 */
    bits_ue(&b, 2); bits_ue(&b, 0);       /* long-term pic 0 = A */
    bits_ue(&b, 2); bits_ue(&b, 1);       /* long-term pic 1 = B */
    if (with_wp)
        for (int i = 0; i < st->n_wp; i++)
            if (st->wp_valid[i]) { bits_ue(&b, 2); bits_ue(&b, (uint32_t)st->wp_lt_idx[i]); }
    bits_ue(&b, 3);
    if (waypoint_frame) {                 /* reference picture: MMCO 4 then 6 (:521-528) */
        int lt = 2 + st->n_wp;
        bits_put(&b, 1, 1);
        bits_ue(&b, 4); bits_ue(&b, (uint32_t)(lt + 1));
        bits_ue(&b, 6); bits_ue(&b, (uint32_t)lt);
        bits_ue(&b, 0);
    }
    bits_se(&b, 0);                       /* slice_qp_delta */
    if (st->deblock_present) bits_ue(&b, 1);
    for (int w = 0; w < 4; w++) {
        uint64_t v = 0;
        for (int k = 0; k < 8; k++) v = (v << 8) | tmp[w * 8 + k];
        hdr[w] = v;
    }
    return (int)b.nbits;
}

int so_needs_waypoint(const SoStream *st, int offset_px) {
    if (offset_px == 0 || offset_px % SO_MV_LIMIT_PX != 0) return 0;
    for (int i = 0; i < st->n_wp; i++)
        if (st->wp_valid[i] && st->wp_offset[i] == offset_px) return 0;
    return 1;
}

void so_register_waypoint(SoStream *st, int offset_px) {
    if (st->n_wp < SO_MAX_WAYPOINTS) {
        st->wp_offset[st->n_wp] = offset_px;
        st->wp_lt_idx[st->n_wp] = 2 + st->n_wp;
        st->wp_valid[st->n_wp] = 1;
        st->n_wp++;
    }
}

void so_legacy_hints(const SoStream *st, int offset_px, int waypoint_frame, H264GpuMbHint *hints) {
    int split = (st->height - offset_px) / 16;                      /* :555 */
    /* A side: nearest waypoint at or below the offset, only beyond the MV limit (:557-571) */
    int ia = -1, oa = 0;
    if (offset_px > SO_MV_LIMIT_PX)
        for (int i = 0; i < st->n_wp; i++) {
            if (!st->wp_valid[i]) continue;
            int wo = st->wp_offset[i];
            if (wo <= offset_px && wo > oa && offset_px - wo <= SO_MV_LIMIT_PX) { ia = i; oa = wo; }
        }
    /* B side: first waypoint above the offset within reach; never in a waypoint frame (:573-588,:726-729) */
    int ib = -1, ob = 0;
    if (!waypoint_frame && offset_px - st->height < -SO_MV_LIMIT_PX)
        for (int i = 0; i < st->n_wp; i++) {
            if (!st->wp_valid[i]) continue;
            int wo = st->wp_offset[i];
            if (wo > offset_px && offset_px - wo >= -SO_MV_LIMIT_PX) { ib = i; ob = wo; break; }
        }
    int ref_a = ia >= 0 ? 2 + ia : 0, mv_a = ia >= 0 ? offset_px - oa : offset_px;
    int ref_b = ib >= 0 ? 2 + ib : 1, mv_b = ib >= 0 ? offset_px - ob : offset_px - st->height;
    for (int y = 0; y < st->mb_h; y++)
        for (int x = 0; x < st->mb_w; x++) {
            H264GpuMbHint *h = &hints[(size_t)y * st->mb_w + x];
            int top = y < split;
            h->mv_x_qpel = 0;
            h->mv_y_qpel = (int16_t)((top ? mv_a : mv_b) * 4);       /* :619-620 */
            h->ref_idx = (uint8_t)(top ? ref_a : ref_b);
            h->kind = H264GPU_MB_INTER;
            h->dyn_idx = 0;
        }
}

size_t so_legacy_stream(int width, int height, int deblock_present,
                        const int *offsets, int n_offsets,
                        uint8_t *out, size_t out_cap,
                        uint32_t *nal_sizes, int *n_nals, int max_nals) {
    SoStream st;
    so_stream_init(&st, width, height, deblock_present);
    H264GpuMbHint *hints = malloc(sizeof(H264GpuMbHint) * (size_t)st.mb_w * st.mb_h);
    size_t pos = 0;
    int k = 0;
    for (int i = 0; i < n_offsets; i++) {
        int wp = so_needs_waypoint(&st, offsets[i]);           /* src/composer.c:257-262 */
        for (int pass = wp ? 0 : 1; pass < 2; pass++) {
            int is_wp = (pass == 0);
            uint64_t hdr[4];
            int nb = so_slice_header(&st, is_wp, hdr);
            so_legacy_hints(&st, offsets[i], is_wp, hints);
            size_t w = so_compose_frame(hints, st.mb_w, st.mb_h, 2 + st.n_wp, 0, hdr, nb,
                                        is_wp ? 0x41 : 0x01, NULL, 0, NULL,
                                        out + pos, out_cap - pos);
            if (is_wp) so_register_waypoint(&st, offsets[i]);
            st.frame_num++;                                    /* :662,:780 */
            pos += w;
            if (k < max_nals) nal_sizes[k] = (uint32_t)w;
            k++;
        }
    }
    free(hints);
    *n_nals = k;
    return pos;
}

double so_time_hinted(const H264GpuMbHint *hints, int n_frames, int mb_w, int mb_h,
                      int num_refs, int flags, const uint64_t *hdr_bits, int hdr_nbits,
                      const uint8_t *payload, size_t payload_bytes,
                      const uint32_t *dyn_bit_off, int tab_stride,
                      int reps, uint64_t *bytes_out) {
    size_t cap = 4u << 20;
    uint8_t *out = malloc(cap);
    struct timespec t0, t1;
    uint64_t total = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++)
        for (int f = 0; f < n_frames; f++)
            total += so_compose_frame(hints + (size_t)f * mb_w * mb_h, mb_w, mb_h, num_refs, flags,
                                      hdr_bits, hdr_nbits, 0x01, payload, payload_bytes,
                                      dyn_bit_off ? dyn_bit_off + (size_t)f * tab_stride : NULL,
                                      out, cap);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(out);
    if (bytes_out) *bytes_out = total;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
