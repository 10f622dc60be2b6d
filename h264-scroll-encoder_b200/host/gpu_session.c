/*
 * gpu_session.c - drives the GPU path for the reference-shaped host API.
 *
 * A legacy scroll frame is at most two horizontal bands of constant motion
 * (src/h264_writer.c:598-617), so the host ships 2 x 16-byte regions, a 64-byte
 * frame descriptor and a 24-byte raster descriptor per frame; the dense hint
 * map, the slice data and the packed Annex-B stream are produced on the device:
 *
 *     H2D(descs) -> raster_regions -> compose_batch -> pack_frames -> D2H(stream)
 *
 * Device and pinned buffers are cached and grown on demand.  Not thread safe
 * per process by itself; callers serialise through the mutex below.
 */
#include "gpu_session.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static struct {
    int ready;
    uint8_t *dev;  size_t dev_cap;
    uint8_t *pin;  size_t pin_cap;
} S;

static size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int session_init(void) {
    if (S.ready) return 0;
    const char *env = getenv("H264GPU_DEVICE");
    int dev = env ? atoi(env) : 0;
    int rc = h264gpu_init(dev);
    if (rc != H264GPU_OK) {
        fprintf(stderr, "Error: GPU composer unavailable (%s); there is no CPU fallback\n",
                h264gpu_last_error());
        return rc;
    }
    S.ready = 1;
    return 0;
}

static int ensure(size_t dev_bytes, size_t pin_bytes) {
    if (dev_bytes > S.dev_cap) {
        h264gpu_free(S.dev);
        S.dev = h264gpu_malloc(dev_bytes);
        S.dev_cap = S.dev ? dev_bytes : 0;
        if (!S.dev) return H264GPU_ERR_NOMEM;
    }
    if (pin_bytes > S.pin_cap) {
        h264gpu_host_free(S.pin);
        S.pin = h264gpu_host_alloc(pin_bytes);
        S.pin_cap = S.pin ? pin_bytes : 0;
        if (!S.pin) return H264GPU_ERR_NOMEM;
    }
    return 0;
}

/* slot bound for a legacy frame: <= 46 bits per MB plus header, 3/2 for worst-case emulation prevention */
static uint32_t legacy_slot_cap(int n_mbs) { return (uint32_t)up(128 + (size_t)n_mbs * 9, 16); }

static int compose_chunk(const LegacyFrame *fr, int n, int mb_w, int mb_h,
                         uint8_t *dst, size_t cap, uint32_t *sizes, size_t *written) {
    const size_t n_mbs = (size_t)mb_w * mb_h;
    const size_t hint_stride = up(n_mbs, 2);
    const uint32_t slot = legacy_slot_cap((int)n_mbs);

    /* pinned staging: descs | rdescs | regions  ...  sizes | offsets */
    size_t p_desc = 0, p_rd = up(p_desc + sizeof(H264GpuFrameDesc) * n, 64);
    size_t p_reg = up(p_rd + sizeof(H264GpuRasterDesc) * n, 64);
    size_t p_up_end = up(p_reg + sizeof(H264GpuRegion) * 2 * n, 64);
    size_t p_sizes = p_up_end, p_offs = up(p_sizes + 4 * (size_t)n, 64);
    size_t p_end = p_offs + 8 * ((size_t)n + 1);

    size_t d_up = 0, d_hints = up(p_up_end, 256);
    size_t d_sizes = up(d_hints + 8 * hint_stride * n, 256);
    size_t d_offs = up(d_sizes + 4 * (size_t)n, 256);
    size_t d_slots = up(d_offs + 8 * ((size_t)n + 1), 256);
    size_t d_packed = up(d_slots + (size_t)slot * n, 256);
    size_t d_end = d_packed + (size_t)slot * n;

    int rc = ensure(d_end, p_end);
    if (rc) return rc;

    H264GpuFrameDesc *desc = (H264GpuFrameDesc *)(S.pin + p_desc);
    H264GpuRasterDesc *rd = (H264GpuRasterDesc *)(S.pin + p_rd);
    H264GpuRegion *reg = (H264GpuRegion *)(S.pin + p_reg);
    for (int f = 0; f < n; f++) {
        const LegacyFrame *L = &fr[f];
        memset(&desc[f], 0, sizeof desc[f]);
        memcpy(desc[f].hdr_bits, L->hdr, sizeof L->hdr);
        desc[f].hint_off = hint_stride * f;
        desc[f].out_off = (uint64_t)slot * f;
        desc[f].out_cap = slot;
        desc[f].dyn_id = H264GPU_NO_DYN;
        desc[f].mb_w = (uint16_t)mb_w;
        desc[f].mb_h = (uint16_t)mb_h;
        desc[f].hdr_nbits = (uint8_t)L->hdr_nbits;
        desc[f].num_refs = (uint8_t)L->num_refs;
        desc[f].flags = H264GPU_SKIP_OFF | H264GPU_PRED_REF;      /* bit-exact reference behaviour */
        desc[f].nal_header = (uint8_t)L->nal_header;
        rd[f].region_first = 2u * f;
        rd[f].region_count = 2;
        rd[f].hint_off = hint_stride * f;
        rd[f].mb_w = (uint16_t)mb_w;
        rd[f].mb_h = (uint16_t)mb_h;
        rd[f].dyn_info_off = 0;
        /* quarter-pel MVs travel as int16: offsets beyond +-8191 px cannot be coded here (the reference's int arithmetic
           would go on; its hardware-decoder limit is +-512 px, include/h264_writer.h:23-24) */
        if (L->mv_top_qpel < -32768 || L->mv_top_qpel > 32767 || L->mv_bot_qpel < -32768 || L->mv_bot_qpel > 32767) return H264GPU_ERR_ARG;
        int split = L->split_row < 0 ? 0 : (L->split_row > mb_h ? mb_h : L->split_row);
        H264GpuRegion *r = &reg[2 * f];
        memset(r, 0, 2 * sizeof *r);
        r[0].mb_x0 = 0; r[0].mb_y0 = 0; r[0].mb_x1 = (int16_t)mb_w; r[0].mb_y1 = (int16_t)split;
        r[0].mv_y_qpel = (int16_t)L->mv_top_qpel; r[0].ref_idx = (uint8_t)L->ref_top;
        r[1].mb_x0 = 0; r[1].mb_y0 = (int16_t)split; r[1].mb_x1 = (int16_t)mb_w; r[1].mb_y1 = (int16_t)mb_h;
        r[1].mv_y_qpel = (int16_t)L->mv_bot_qpel; r[1].ref_idx = (uint8_t)L->ref_bot;
    }

    uint8_t *D = S.dev;
    rc = h264gpu_memcpy_h2d(D + d_up, S.pin, p_up_end, NULL);
    if (!rc) rc = h264gpu_raster_regions((const H264GpuRasterDesc *)(D + p_rd), n, (int)n_mbs,
                                         (const H264GpuRegion *)(D + p_reg), NULL,
                                         (H264GpuMbHint *)(D + d_hints), NULL);
    if (!rc) rc = h264gpu_compose_batch_ex((const H264GpuFrameDesc *)(D + p_desc), n,
                                        (const H264GpuMbHint *)(D + d_hints), hint_stride * (size_t)n, NULL, NULL, NULL,
                                        D + d_slots, (uint32_t *)(D + d_sizes), NULL);
    if (!rc) rc = h264gpu_pack_frames((const H264GpuFrameDesc *)(D + p_desc), (const uint32_t *)(D + d_sizes), n,
                                      D + d_slots, D + d_packed, (uint64_t *)(D + d_offs), NULL);
    if (!rc) rc = h264gpu_memcpy_d2h(S.pin + p_sizes, D + d_sizes, 4 * (size_t)n, NULL);
    if (!rc) rc = h264gpu_memcpy_d2h(S.pin + p_offs, D + d_offs, 8 * ((size_t)n + 1), NULL);
    if (!rc) rc = h264gpu_stream_sync(NULL);
    if (rc) return rc;

    const uint32_t *hs = (const uint32_t *)(S.pin + p_sizes);
    const uint64_t *ho = (const uint64_t *)(S.pin + p_offs);
    for (int f = 0; f < n; f++) {
        if (hs[f] > slot) return H264GPU_ERR_OVERFLOW;
        sizes[f] = hs[f];
    }
    size_t total = (size_t)ho[n];
    if (total > cap) return H264GPU_ERR_OVERFLOW;
    rc = h264gpu_memcpy_d2h(dst, D + d_packed, total, NULL);
    if (!rc) rc = h264gpu_stream_sync(NULL);
    if (rc) return rc;
    *written = total;
    return 0;
}

/*
 * One frame per call - what composer_write_scroll_frame does (src/composer.c:255-264) - is latency, not throughput:
 * the reference takes ~144 us for a 1280x720 frame.  So no staging copies and no packing here: the three small
 * descriptors sit in pinned host memory and the kernels read them from there, the compose kernel stores the frame's
 * bytes and its size straight into pinned host memory, and the host waits once.
 *     raster_regions (descriptors over PCIe -> dense map in HBM) -> compose_batch (Annex-B bytes over PCIe) -> sync
 */
static int compose_one(const LegacyFrame *L, int mb_w, int mb_h, uint8_t *dst, size_t cap, uint32_t *size_out, size_t *written) {
    const size_t n_mbs = (size_t)mb_w * mb_h;
    const uint32_t slot = legacy_slot_cap((int)n_mbs);
    /* pinned: desc | rdesc | 2 regions | size | slot */
    const size_t p_desc = 0, p_rd = 64, p_reg = 128, p_size = 192, p_slot = 256, p_end = p_slot + slot;
    int rc = ensure(8 * up(n_mbs, 2), p_end);
    if (rc) return rc;
    if (L->mv_top_qpel < -32768 || L->mv_top_qpel > 32767 || L->mv_bot_qpel < -32768 || L->mv_bot_qpel > 32767) return H264GPU_ERR_ARG;
    H264GpuFrameDesc *desc = (H264GpuFrameDesc *)(S.pin + p_desc);
    H264GpuRasterDesc *rd = (H264GpuRasterDesc *)(S.pin + p_rd);
    H264GpuRegion *r = (H264GpuRegion *)(S.pin + p_reg);
    volatile uint32_t *h_size = (volatile uint32_t *)(S.pin + p_size);
    memset(desc, 0, sizeof *desc);
    memcpy(desc->hdr_bits, L->hdr, sizeof L->hdr);
    desc->out_cap = slot;
    desc->dyn_id = H264GPU_NO_DYN;
    desc->mb_w = (uint16_t)mb_w; desc->mb_h = (uint16_t)mb_h;
    desc->hdr_nbits = (uint8_t)L->hdr_nbits;
    desc->num_refs = (uint8_t)L->num_refs;
    desc->flags = H264GPU_SKIP_OFF | H264GPU_PRED_REF;
    desc->nal_header = (uint8_t)L->nal_header;
    memset(rd, 0, sizeof *rd);
    rd->region_count = 2;
    rd->mb_w = (uint16_t)mb_w; rd->mb_h = (uint16_t)mb_h;
    const int split = L->split_row < 0 ? 0 : (L->split_row > mb_h ? mb_h : L->split_row);
    memset(r, 0, 2 * sizeof *r);
    r[0].mb_x1 = (int16_t)mb_w; r[0].mb_y1 = (int16_t)split;
    r[0].mv_y_qpel = (int16_t)L->mv_top_qpel; r[0].ref_idx = (uint8_t)L->ref_top;
    r[1].mb_y0 = (int16_t)split; r[1].mb_x1 = (int16_t)mb_w; r[1].mb_y1 = (int16_t)mb_h;
    r[1].mv_y_qpel = (int16_t)L->mv_bot_qpel; r[1].ref_idx = (uint8_t)L->ref_bot;
    *h_size = 0;
    rc = h264gpu_raster_regions(rd, 1, (int)n_mbs, r, NULL, (H264GpuMbHint *)S.dev, NULL);
    if (!rc) rc = h264gpu_compose_batch_ex(desc, 1, (const H264GpuMbHint *)S.dev, up(n_mbs, 2), NULL, NULL, NULL, S.pin + p_slot, (uint32_t *)(S.pin + p_size), NULL);
    if (!rc) rc = h264gpu_stream_sync(NULL);
    if (rc) return rc;
    const uint32_t sz = *h_size;
    if (sz > slot || sz > cap) return H264GPU_ERR_OVERFLOW;
    memcpy(dst, S.pin + p_slot, sz);
    *size_out = sz;
    *written = sz;
    return 0;
}

int gpu_session_compose_legacy(const LegacyFrame *frames, int n, int mb_w, int mb_h,
                               uint8_t *dst, size_t cap, uint32_t *sizes, size_t *total) {
    if (n <= 0) { *total = 0; return 0; }
    if (mb_w <= 0 || mb_h <= 0 || mb_w > H264GPU_MAX_MB_W) return H264GPU_ERR_ARG;
    pthread_mutex_lock(&g_lock);
    int rc = session_init();
    size_t done_bytes = 0;
    if (!rc && n == 1) {
        rc = compose_one(frames, mb_w, mb_h, dst, cap, sizes, &done_bytes);
    } else if (!rc) {
        /* bound device scratch to ~1 GiB per launch */
        size_t per_frame = (size_t)mb_w * mb_h * 8 + 2 * (size_t)legacy_slot_cap(mb_w * mb_h) + 256;
        int chunk = (int)((1ull << 30) / per_frame);
        if (chunk < 1) chunk = 1;
        for (int f0 = 0; f0 < n && !rc; f0 += chunk) {
            int m = n - f0 < chunk ? n - f0 : chunk;
            size_t w = 0;
            rc = compose_chunk(frames + f0, m, mb_w, mb_h, dst + done_bytes, cap - done_bytes, sizes + f0, &w);
            done_bytes += w;
        }
    }
    pthread_mutex_unlock(&g_lock);
    if (rc && rc != H264GPU_ERR_NO_DEVICE)
        fprintf(stderr, "Error: GPU compose failed (%d: %s)\n", rc, h264gpu_last_error());
    *total = done_bytes;
    return rc;
}

/* ------------------------------------------------------------ hinted frames */

#include "composer.h"

/* a frame with hints: worst-case inline code 104 bits per macroblock, payload once, 3/2 for emulation prevention */
static uint32_t hinted_slot_cap(size_t n_mbs, size_t payload_bytes) {
    return (uint32_t)up(256 + n_mbs * 20 + payload_bytes * 3 / 2, 16);
}

int gpu_session_compose_hinted(const LegacyFrame *fr, const HintOverlay *ov, int n, int mb_w, int mb_h,
                               uint8_t *dst, size_t cap, uint32_t *sizes, size_t *total) {
    *total = 0;
    if (n <= 0) return 0;
    if (mb_w <= 0 || mb_h <= 0 || mb_w > H264GPU_MAX_MB_W) return H264GPU_ERR_ARG;
    const size_t n_mbs = (size_t)mb_w * mb_h, hint_stride = up(n_mbs, 2);

    /* sizes of the batch's arrays */
    size_t n_reg = 0, n_info = 0, n_tab = 0, n_pay = 0;
    int n_dyn = 0;
    for (int f = 0; f < n; f++) {
        n_reg += 2;
        if (!ov || (!ov[f].regions && !ov[f].dyn)) continue;
        if (ov[f].n_regions < 0 || ov[f].n_regions + 3 > H264GPU_MAX_REGIONS) return H264GPU_ERR_ARG;
        n_reg += (size_t)ov[f].n_regions;
        const ComposerDynRegion *d = ov[f].dyn;
        if (d) {
            if (d->mb_x0 < 0 || d->mb_y0 < 0 || d->mb_x1 > mb_w || d->mb_y1 > mb_h || d->mb_x1 <= d->mb_x0 || d->mb_y1 <= d->mb_y0)
                return H264GPU_ERR_ARG;                          /* the rectangle must lie inside the frame */
            const size_t m = (size_t)(d->mb_x1 - d->mb_x0) * (size_t)(d->mb_y1 - d->mb_y0);
            if (m > 65535 || !d->mb_bit_off || !d->mb_info || (d->payload_bytes && !d->payload)) return H264GPU_ERR_ARG;
            if ((size_t)d->mb_bit_off[m] > 8 * d->payload_bytes) return H264GPU_ERR_ARG;
            n_reg += 1; n_info += m; n_tab += m + 1; n_pay += up(d->payload_bytes, 4); n_dyn++;
        }
    }
    H264GpuFrameDesc *desc = calloc((size_t)n, sizeof *desc);
    H264GpuRasterDesc *rd = calloc((size_t)n, sizeof *rd);
    H264GpuRegion *reg = calloc(n_reg, sizeof *reg);
    H264GpuMbHint *info = calloc(n_info ? n_info : 1, sizeof *info);
    H264GpuDynDesc *dyn = calloc((size_t)(n_dyn ? n_dyn : 1), sizeof *dyn);
    uint32_t *tab = calloc(n_tab ? n_tab : 1, sizeof *tab);
    uint8_t *pay = calloc(n_pay + 4, 1);
    uint64_t *offs = calloc((size_t)n + 1, sizeof *offs);
    int rc = (desc && rd && reg && info && dyn && tab && pay && offs) ? 0 : H264GPU_ERR_NOMEM;

    size_t r = 0, i0 = 0, t0 = 0, p0 = 0, slots = 0;
    int y = 0;
    for (int f = 0; f < n && !rc; f++) {
        const LegacyFrame *L = &fr[f];
        const ComposerDynRegion *d = ov ? ov[f].dyn : NULL;
        const uint32_t slot = hinted_slot_cap(n_mbs, d ? d->payload_bytes : 0);
        memcpy(desc[f].hdr_bits, L->hdr, sizeof L->hdr);
        desc[f].hint_off = hint_stride * (size_t)f;
        desc[f].out_off = slots;
        desc[f].out_cap = slot;
        slots += slot;
        desc[f].dyn_id = H264GPU_NO_DYN;
        desc[f].mb_w = (uint16_t)mb_w; desc[f].mb_h = (uint16_t)mb_h;
        desc[f].hdr_nbits = (uint8_t)L->hdr_nbits;
        desc[f].num_refs = (uint8_t)L->num_refs;
        desc[f].flags = H264GPU_SKIP_OFF | H264GPU_PRED_REF;      /* the reference's coding decisions */
        desc[f].nal_header = (uint8_t)L->nal_header;
        rd[f].region_first = (uint32_t)r;
        rd[f].hint_off = desc[f].hint_off;
        rd[f].mb_w = (uint16_t)mb_w; rd[f].mb_h = (uint16_t)mb_h;
        rd[f].dyn_info_off = (uint32_t)i0;
        /* the scroll frame's two bands (src/h264_writer.c:598-617) ... */
        if (L->mv_top_qpel < -32768 || L->mv_top_qpel > 32767 || L->mv_bot_qpel < -32768 || L->mv_bot_qpel > 32767) { rc = H264GPU_ERR_ARG; break; }
        int split = L->split_row < 0 ? 0 : (L->split_row > mb_h ? mb_h : L->split_row);
        reg[r].mb_x1 = (int16_t)mb_w; reg[r].mb_y1 = (int16_t)split;
        reg[r].mv_y_qpel = (int16_t)L->mv_top_qpel; reg[r].ref_idx = (uint8_t)L->ref_top; r++;
        reg[r].mb_y0 = (int16_t)split; reg[r].mb_x1 = (int16_t)mb_w; reg[r].mb_y1 = (int16_t)mb_h;
        reg[r].mv_y_qpel = (int16_t)L->mv_bot_qpel; reg[r].ref_idx = (uint8_t)L->ref_bot; r++;
        /* ... the caller's motion rectangles over them, in list order ... */
        const ComposerMotionRegion *mr = ov ? (const ComposerMotionRegion *)ov[f].regions : NULL;
        for (int k = 0; mr && k < ov[f].n_regions; k++) {
            if (mr[k].mv_x_px < -8192 || mr[k].mv_x_px > 8191 || mr[k].mv_y_px < -8192 || mr[k].mv_y_px > 8191 || mr[k].ref_idx >= L->num_refs) { rc = H264GPU_ERR_ARG; break; }
            reg[r].mb_x0 = mr[k].mb_x0; reg[r].mb_y0 = mr[k].mb_y0; reg[r].mb_x1 = mr[k].mb_x1; reg[r].mb_y1 = mr[k].mb_y1;
            reg[r].mv_x_qpel = (int16_t)(4 * mr[k].mv_x_px); reg[r].mv_y_qpel = (int16_t)(4 * mr[k].mv_y_px);
            reg[r].ref_idx = mr[k].ref_idx; r++;
        }
        /* ... and the dynamic rectangle on top of everything */
        if (d && !rc) {
            const size_t m = (size_t)(d->mb_x1 - d->mb_x0) * (size_t)(d->mb_y1 - d->mb_y0);
            reg[r].mb_x0 = d->mb_x0; reg[r].mb_y0 = d->mb_y0; reg[r].mb_x1 = d->mb_x1; reg[r].mb_y1 = d->mb_y1;
            reg[r].kind = H264GPU_REGION_DYN; reg[r].dyn_base = 0; r++;
            for (size_t j = 0; j < m; j++) {
                const ComposerMbInfo *q = &d->mb_info[j];
                info[i0 + j].mv_x_qpel = q->intra ? 0 : q->mv_x_qpel;
                info[i0 + j].mv_y_qpel = q->intra ? 0 : q->mv_y_qpel;
                info[i0 + j].ref_idx = q->intra ? 0 : q->ref_idx;
                info[i0 + j].kind = q->intra ? H264GPU_MB_DYN_INTRA : H264GPU_MB_DYN_INTER;
                info[i0 + j].dyn_idx = (uint16_t)j;
                if (d->mb_bit_off[j + 1] < d->mb_bit_off[j]) rc = H264GPU_ERR_ARG;
            }
            memcpy(tab + t0, d->mb_bit_off, (m + 1) * sizeof *tab);
            if (d->payload_bytes) memcpy(pay + p0, d->payload, d->payload_bytes);
            dyn[y].payload_off = p0; dyn[y].table_off = t0; dyn[y].n_mbs = (uint32_t)m;
            dyn[y].payload_bytes = (uint32_t)up(d->payload_bytes, 4);
            desc[f].dyn_id = (uint32_t)y;
            i0 += m; t0 += m + 1; p0 += up(d->payload_bytes, 4); y++;
        }
        rd[f].region_count = (uint32_t)(r - rd[f].region_first);
    }
    if (!rc) {
        H264GpuHostBatch b;
        memset(&b, 0, sizeof b);
        b.desc = desc; b.n_frames = n; b.rdesc = rd; b.regions = reg; b.n_regions = r;
        b.dyn_info = n_info ? info : NULL; b.n_dyn_info = n_info;
        b.dyn = n_dyn ? dyn : NULL; b.n_dyn = n_dyn;
        b.dyn_tab = n_tab ? tab : NULL; b.n_tab = n_tab;
        b.payload = n_pay ? pay : NULL; b.payload_bytes = n_pay;
        b.n_hints = hint_stride * (size_t)n; b.slots_bytes = slots; b.max_mbs_per_frame = (int)n_mbs;
        pthread_mutex_lock(&g_lock);
        rc = session_init();
        if (!rc) rc = h264gpu_compose_regions_host(&b, dst, cap, offs, sizes, 0);
        pthread_mutex_unlock(&g_lock);
        if (!rc) *total = (size_t)offs[n];
        else if (rc != H264GPU_ERR_NO_DEVICE) fprintf(stderr, "Error: GPU compose failed (%d: %s)\n", rc, h264gpu_last_error());
    }
    free(desc); free(rd); free(reg); free(info); free(dyn); free(tab); free(pay); free(offs);
    return rc;
}
