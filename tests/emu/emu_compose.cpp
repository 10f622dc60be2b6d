/*
 * emu_compose.cpp - TEST ONLY.  Sequential CPU emulation of the tile pipeline of
 * csrc/compose_kernels.cu (load -> decide -> skiprun -> scan -> pack -> epb),
 * built on the very same mb_codec.cuh helpers the kernel uses.  It lets the
 * CPU test-suite check the bit-level logic (entry layout, word-owned gather,
 * chunking, closed-form emulation prevention) against the oracle without a GPU,
 * and with tiny TILE / CH values that force every boundary case.
 *
 * Not part of the product: nothing outside tests/ builds or loads this file.
 */
#include <vector>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include "mb_codec.cuh"

using namespace h264scroll;

struct H2 { u32 x, y; };

extern "C" uint32_t emu_compose_frame(const H264GpuFrameDesc *desc, const H264GpuMbHint *hints_base,
                                      const H264GpuDynDesc *dyns, const uint32_t *dyn_tab,
                                      const uint8_t *payload, uint8_t *out_base,
                                      int TILE, int CH) {
    const H264GpuFrameDesc &d = *desc;
    const int W = d.mb_w, N = W * (int)d.mb_h;
    const int num_refs = d.num_refs;
    const int skip_spec = (d.flags & H264GPU_SKIP_SPEC) != 0;
    const int pred_spec = (d.flags & H264GPU_PRED_SPEC) != 0;
    const u32 cap = d.out_cap;
    const H2 *fh = (const H2 *)hints_base + d.hint_off;
    const u32 *tab = nullptr, *pay_words = nullptr;
    u32 n_pay_words = 0;
    if (d.dyn_id != H264GPU_NO_DYN) {
        const H264GpuDynDesc &dd = dyns[d.dyn_id];
        tab = dyn_tab + dd.table_off;
        pay_words = (const u32 *)(payload + dd.payload_off);
        n_pay_words = (dd.payload_bytes + 3u) >> 2;
    }
    uint8_t *fout = out_base + d.out_off;

    u32 bitpos = 40u + d.hdr_nbits, carry = 0, ins_total = 0;
    int last_coded = -1, zc = 0;
    u32 size = 0;

    std::vector<MbEntry> ent(TILE + 1);
    std::vector<u32> off(TILE + 4);
    std::vector<int> codedv(TILE);
    std::vector<u32> kinds(TILE + 2, KIND_SHORT);
    std::vector<u32> outw(CH);
    std::vector<int> first(CH);

    for (int t0 = 0; t0 < N; t0 += TILE) {
        const int n_real = std::min(TILE, N - t0);
        const bool last = (t0 + TILE >= N), tile0 = (t0 == 0);

        /* decide + code (neighbours read straight from the frame's hint map) */
        for (int i = 0; i < TILE; i++) {
            const int g = t0 + i;
            int coded = 0;
            u32 len = 0, kind = KIND_SHORT;
            if (i < n_real) {
                const H2 h = fh[g];
                const u32 hk = (h.y >> 8) & 0xFFu;
                MbEntry e; e.w0 = e.w1 = e.w2 = e.w3 = 0;
                coded = 1;
                if (hk == H264GPU_MB_INTER) {
                    const int y = g / W, x = g - y * W;
                    const int ref = (int)(h.y & 0xFFu);
                    Nbr a = {0, 0, 0}, b = {0, 0, 0}, c = {0, 0, 0};
                    const int avA = x > 0, avB = y > 0;
                    int avC = 0;
                    if (avA) a = nbr_from_hint(fh[g - 1].x, fh[g - 1].y);
                    if (avB) {
                        b = nbr_from_hint(fh[g - W].x, fh[g - W].y);
                        if (x + 1 < W) { c = nbr_from_hint(fh[g - W + 1].x, fh[g - W + 1].y); avC = 1; }
                        else if (x > 0) { c = nbr_from_hint(fh[g - W - 1].x, fh[g - W - 1].y); avC = 1; }
                    }
                    const int mvx = (int)(int16_t)(h.x & 0xFFFFu), mvy = (int)(int16_t)(h.x >> 16);
                    int px, py;
                    /* the kernel's uniform-neighbourhood shortcut: A = B = C = own hint */
                    const int cidx = x + 1 < W ? g - W + 1 : g - W - 1;
                    const bool uniform = avA && avB && avC &&
                        fh[g - 1].x == h.x && fh[g - W].x == h.x && fh[cidx].x == h.x &&
                        (fh[g - 1].y & 0xFFFFu) == (h.y & 0xFFFFu) && (fh[g - W].y & 0xFFFFu) == (h.y & 0xFFFFu) &&
                        (fh[cidx].y & 0xFFFFu) == (h.y & 0xFFFFu);
                    if (uniform) {
                        px = mvx; py = mvy;
                        if (skip_spec && ref == 0) coded = 0;
                    } else {
                        predict_mv(pred_spec, a, avA, b, avB, c, avC, ref, px, py);
                        if (skip_spec && ref == 0) {
                            int sx, sy;
                            skip_infer_mv(pred_spec, a, avA, b, avB, c, avC, sx, sy);
                            if (mvx == sx && mvy == sy) coded = 0;
                        }
                    }
                    if (coded) {
                        int blen = encode_p16x16_body(ref, mvx - px, mvy - py, num_refs, e);
                        len = (u32)blen;
                        if (!skip_spec) { kind = finish_inline_entry(e, 0x80000000u, 1, blen); len += 1u; }
                    }
                } else {
                    const u32 di = h.y >> 16;
                    e.w1 = tab[di]; e.w2 = tab[di + 1] - tab[di]; e.w3 = ENT_DYN;
                    len = e.w2;
                    kind = KIND_DYN;
                    if (!skip_spec) { e.w0 = 0x80000000u; e.w3 |= 1u; len += 1u; }
                }
                ent[i] = e;
            }
            off[i] = len;
            kinds[i] = kind;
            codedv[i] = coded;
        }
        /* mb_skip_run */
        int last_coded_next = last_coded;
        if (skip_spec) {
            int prev = last_coded;
            for (int i = 0; i < TILE; i++) {
                if (!codedv[i]) continue;
                const u32 run = (u32)(t0 + i - 1 - prev);
                u32 bits; int n;
                encode_skip_prefix(run, bits, n);
                if (kinds[i] == KIND_DYN) { ent[i].w0 = bits; ent[i].w3 |= (u32)n; }
                else kinds[i] = finish_inline_entry(ent[i], bits, n, (int)off[i]);
                off[i] += (u32)n;
                prev = t0 + i;
            }
            last_coded_next = prev;
        }
        u32 trailer_len = 0;
        if (last) {
            u32 bits = 0x80000000u; int n = 1;
            if (skip_spec) {
                const u32 run_end = (u32)(N - 1 - last_coded_next);
                if (run_end > 0) { encode_skip_prefix(run_end, bits, n); bits |= 1u << (31 - n); n += 1; }
            }
            MbEntry e; e.w0 = bits; e.w1 = 0; e.w2 = 0; e.w3 = 0;
            ent[n_real] = e;
            kinds[n_real] = KIND_SHORT;
            trailer_len = (u32)n;
        }
        /* scan */
        {
            u32 run = bitpos;
            for (int i = 0; i < TILE; i++) { u32 l = off[i]; off[i] = run; run += l; }
            off[n_real] = run;
            if (last) off[n_real + 1] = run + trailer_len;
        }
        const int n = last ? n_real + 1 : n_real;
        const u32 end = off[n];
        const u32 w_lo = bitpos >> 5;
        const u32 total_bytes = (end + 7u) >> 3;
        const u32 w_end = end >> 5;
        const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;
        const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end;
        const u32 cb0 = tile0 ? 0u : w_lo;
        u32 carry_next = 0;

        for (u32 cb = cb0; cb <= w_last; cb += (u32)CH) {
            std::fill(first.begin(), first.end(), -1);
            for (int i = 0; i <= n; i++) {
                const u32 j = off[i] >> 5;
                const bool fiw = (i == 0) || ((off[i - 1] >> 5) != j);
                if (fiw && j >= cb && j < cb + (u32)CH) first[j - cb] = i;
            }
            if (tile0) for (u32 t = 0; t < (u32)CH && cb + t < w_lo; t++) first[t] = 0;
            const u32 j_hi = std::min(cb + (u32)CH, w_last + 1u);
            for (u32 j = cb; j < j_hi; j++) {
                if (first[j - cb] < 0) continue;
                const u32 base = j << 5;
                u32 acc = tile0 ? (base < 296u ? header_bits32(d.hdr_bits, d.nal_header, base) : 0u)
                                : (j == w_lo ? carry : 0u);
                int i = first[j - cb];
                auto contrib = [&](int e) {
                    return kinds[e] == KIND_SHORT ? short_contrib(ent[e].w0, off[e], base)
                                                  : entry_contrib(ent[e], off[e], base, pay_words, n_pay_words);
                };
                if (i > 0) acc |= contrib(i - 1);
                while (i < n && off[i] < base + 32u) { acc |= contrib(i); i++; }
                if (j < w_stop) outw[j - cb] = acc; else carry_next = acc;
            }
            for (int i = 0; i < n; i++) {
                if (kinds[i] == KIND_SHORT) continue;
                u32 ja = (off[i] >> 5) + 1u, jb = off[i + 1] >> 5;
                if (ja < cb) ja = cb;
                if (jb > cb + (u32)CH) jb = cb + (u32)CH;
                for (u32 j = ja; j < jb; j++) outw[j - cb] = entry_contrib(ent[i], off[i], j << 5, pay_words, n_pay_words);
            }
            const u32 e_hi = std::min(cb + (u32)CH, w_stop);
            const int cw = e_hi > cb ? (int)(e_hi - cb) : 0;
            const u32 chunk_byte0 = cb << 2;
            const int chunk_bytes = last ? (int)(std::min(e_hi << 2, total_bytes) - chunk_byte0) : cw * 4;
            std::vector<u32> masks(cw, 0);
            bool any = false;
            int prev_nz = -1 - zc;
            for (int jj = 0; jj < cw; jj++) {
                int vb = std::min(4, chunk_bytes - 4 * jj);
                masks[jj] = epb_word_mask(outw[jj], vb, 4 * jj, chunk_byte0, prev_nz);
                any = any || masks[jj];
            }
            const int r_end = chunk_bytes - 1 - prev_nz;
            const int zc_next = r_end == 0 ? 0 : ((r_end & 1) ? 1 : 2);
            if (!any && ins_total == 0) {
                for (int jj = 0; jj < cw; jj++) {
                    const u32 j = cb + jj;
                    if (4u * j + 4u <= cap) { u32 v = bswap32(outw[jj]); memcpy(fout + 4 * j, &v, 4); }
                }
            } else {
                u32 running = ins_total;
                for (int jj = 0; jj < cw; jj++) {
                    const int vb = std::min(4, chunk_bytes - 4 * jj);
                    u32 pos = chunk_byte0 + 4u * jj + running;
                    for (int k = 0; k < vb; k++) {
                        if ((masks[jj] >> k) & 1u) { if (pos < cap) fout[pos] = 3; pos++; }
                        if (pos < cap) fout[pos] = (uint8_t)chunk_byte(outw.data(), 4 * jj + k);
                        pos++;
                    }
                    running += (u32)__builtin_popcount(masks[jj]);
                }
                ins_total = running;
            }
            if (cw > 0) zc = zc_next;
        }
        bitpos = end;
        carry = carry_next;
        last_coded = last_coded_next;
        if (last) size = total_bytes + ins_total;
    }
    return size;
}
