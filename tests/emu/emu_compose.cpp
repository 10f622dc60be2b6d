/*
 * emu_compose.cpp - TEST ONLY.  Sequential CPU emulation of the tile pipeline of
 * csrc/compose_kernels.cu (segments -> decide -> skiprun -> scan -> concatenate/place -> queue items ->
 * emit with pre-flush / overflow passes), built on the very same mb_codec.cuh helpers the kernel
 * uses (decide_mb, attach_runs, concat_append/_place, make_item, consume_item, epb_group_mask,
 * realign_selector ...).  "Threads" are walked one after the other; the warp-shuffle merge of partial
 * words becomes a plain OR into the window.  It lets the CPU test-suite check the bit-level logic
 * against the oracle without a GPU, with tiny thread counts / window sizes that force every boundary case.
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
static const int MPT = 8;

static u32 byte_perm(u32 a, u32 b, u32 sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    u32 r = 0;
    for (int k = 0; k < 4; k++) r |= (u32)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
}

template <int NWc>
static void simple_words_emu(const u32 (&code)[8], const u32 (&len)[8], u32 O, u32 Ls, u32 *win, u32 wbase, u32 CH) {
    u32 a[NWc];
    for (int i = 0; i < NWc; i++) a[i] = 0;
    for (int k = 0; k < 8; k++) concat_append<NWc>(a, code[k], len[k], k);
    const u32 E = O + Ls, jl = (E - 1u) >> 5, j0 = O >> 5;
    u32 b[NWc + 1];
    concat_place<NWc>(a, (32u - (E & 31u)) & 31u, b);
    for (int i = 0; i <= NWc; i++) {
        const u32 idx = jl - (u32)(NWc - i);
        if ((int)(idx - j0) >= 0 && idx - wbase < CH) win[idx - wbase] |= b[i];
        else if ((int)(idx - j0) < 0 && b[i] != 0) __builtin_trap();      /* words above the string must be empty */
    }
}

extern "C" uint32_t emu_compose_frame(const H264GpuFrameDesc *desc, const H264GpuMbHint *hints_base,
                                      const H264GpuDynDesc *dyns, const uint32_t *dyn_tab,
                                      const uint8_t *payload, uint8_t *out_base,
                                      int NTHREADS, int CHi) {
    const H264GpuFrameDesc &d = *desc;
    const u32 CH = (u32)CHi;
    const int W = d.mb_w, Hh = d.mb_h;
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
    const u32 hdr_total = 40u + d.hdr_nbits;
    const int tpr = (W + MPT - 1) / MPT, n_seg = tpr * Hh;
    const int seg_max = std::max(1, std::min(NTHREADS, NTHREADS + 34 - tpr - 1));
    const int kc = (W - 1) & 7;

    u32 bitpos = hdr_total, ins_total = 0, zrun = 0, wbase = 0;
    int last_coded = -1;
    u32 size = 0;
    const H2 unavail = {0u, HINT_Y_UNAVAIL};
    auto HN = [&](int row, int x) -> H2 {              /* hint at (row, x) in canonical form, "unavailable" outside */
        if (row < 0 || x < 0 || x >= W) return unavail;
        const H2 h = fh[row * W + x];
        return H2{canonical_hint_x(h.x, h.y), canonical_hint_y(h.y)};
    };

    std::vector<u32> win(CH + 8, 0u);

    /* emit of the window's first chunk_bytes bytes (frame bytes from chunk_byte0 on), as emit_window in the kernel */
    auto emit_window = [&](u32 chunk_byte0, int chunk_bytes) {
        const u32 zrun_in = zrun;
        for (int q0 = 0; q0 * 16 < chunk_bytes; q0 += NTHREADS) {
            std::vector<u32> insv(NTHREADS, 0);
            std::vector<int> vbs(NTHREADS, 0);
            for (int t = 0; t < NTHREADS; t++) {
                const int q = q0 + t;
                const int vb = std::min(std::max(chunk_bytes - 16 * q, 0), 16);
                vbs[t] = vb;
                if (vb == 0) continue;
                u32 w[4];
                for (int i = 0; i < 4; i++) w[i] = win[4 * q + i];
                const u32 hz = ((w[0] - 0x01010101u) & ~w[0]) | ((w[1] - 0x01010101u) & ~w[1]) |
                               ((w[2] - 0x01010101u) & ~w[2]) | ((w[3] - 0x01010101u) & ~w[3]);
                if ((hz & 0x80808080u) || (w[0] >> 26) == 0) {
                    const u32 valid = (1u << vb) - 1u;
                    u32 z = (zero_nib(w[0]) | (zero_nib(w[1]) << 4) | (zero_nib(w[2]) << 8) | (zero_nib(w[3]) << 12)) & valid;
                    u32 keep = valid;
                    if (chunk_byte0 == 0 && q == 0) { z &= ~0x1Fu; keep &= ~0x1Fu; }
                    const bool need = ((w[0] >> 26) == 0) && (keep & 1);
                    if (need || (z & (z << 1))) {
                        u32 rin = 0;
                        if (need && (q == 0 || (win[4 * q - 1] & 0xFF) == 0)) rin = zeros_before(win.data(), 16u * q, zrun_in);
                        const u32 m3 = zero_nib(w[0] & 0xFCFCFCFCu) | (zero_nib(w[1] & 0xFCFCFCFCu) << 4) |
                                       (zero_nib(w[2] & 0xFCFCFCFCu) << 8) | (zero_nib(w[3] & 0xFCFCFCFCu) << 12);
                        insv[t] = epb_group_mask(z, m3 & keep, vb, rin);
                    }
                }
                if (16 * (q + 1) >= chunk_bytes) zrun = zeros_before(win.data(), (u32)chunk_bytes, zrun_in);
            }
            u32 pass_ins = 0;
            for (int t = 0; t < NTHREADS; t++) pass_ins += (u32)__builtin_popcount(insv[t]);
            const bool roomy = chunk_byte0 + 16u * (u32)(q0 + NTHREADS) + ins_total + pass_ins + 4u <= cap;
            u32 before = ins_total;
            for (int t = 0; t < NTHREADS; t++) {
                const int q = q0 + t, vb = vbs[t];
                if (vb == 0) continue;
                const u32 cnt = (u32)__builtin_popcount(insv[t]);
                const u32 D = chunk_byte0 + 16u * q + before;
                auto full_clean = [&](int u) { return u >= 0 && u < NTHREADS && vbs[u] == 16 && insv[u] == 0 && roomy; };
                const bool clean = cnt == 0 && (vb & 3) == 0 && roomy;
                const bool prev_clean = full_clean(t - 1), next_clean = full_clean(t + 1);
                u32 w[5];
                for (int i = 0; i < 5; i++) w[i] = win[4 * q + i];
                if (clean) {
                    const u32 s = (4u - (D & 3u)) & 3u;
                    const u32 sel = realign_selector(s);
                    const int nw = vb >> 2;
                    u32 o[4];
                    for (int i = 0; i < 4; i++) o[i] = byte_perm(w[i], w[i + 1], sel);
                    const bool whole_last = s == 0 || (vb == 16 && next_clean);
                    for (int i = 0; i < nw; i++)
                        if (i < nw - 1 || whole_last) memcpy(fout + D + s + 4 * i, &o[i], 4);
                    if (!whole_last)
                        for (u32 k = 0; k < 4u - s; k++) fout[D + s + 4 * (nw - 1) + k] = (uint8_t)(o[nw - 1] >> (8 * k));
                    if (s && (!prev_clean || vb < 16))
                        for (u32 k = 0; k < s; k++) fout[D + k] = (uint8_t)(w[0] >> (24 - 8 * k));
                } else {
                    u32 pos = D;
                    for (int k = 0; k < vb; k++) {
                        if ((insv[t] >> k) & 1) { if (pos < cap) fout[pos] = 3; pos++; }
                        if (pos < cap) fout[pos] = (uint8_t)(w[k >> 2] >> (24 - 8 * (k & 3)));
                        pos++;
                    }
                }
                before += cnt;
            }
            ins_total += pass_ins;
        }
    };

    struct Th { u32 code[8], len[8], kinds, Ls, L, O; int g0, prev0; bool cplx, trailer; H2 own[8]; };
    std::vector<Th> th(NTHREADS);

    for (int s0 = 0; s0 < n_seg;) {
        const int n = std::min(seg_max, n_seg - s0);
        const bool last = (s0 + n >= n_seg), tile0 = (s0 == 0);
        /* decide */
        for (int t = 0; t < n; t++) {
            Th &T = th[t];
            const int s = s0 + t, row = s / tpr, col = s - row * tpr;
            const int nmb = std::min(MPT, W - MPT * col), x0 = MPT * col;
            const bool lastseg = col == tpr - 1;
            T.g0 = row * W + x0;
            T.kinds = 0;
            H2 own[MPT + 1], abv[MPT + 2];
            for (int k = 0; k <= MPT; k++) own[k] = HN(row, x0 - 1 + k);
            for (int j = 0; j < MPT + 2; j++) abv[j] = HN(row - 1, x0 - 1 + j);
            if (lastseg) abv[kc + 2] = abv[kc];
            for (int k = 0; k < MPT; k++) {
                T.code[k] = 0; T.len[k] = 0;
                T.own[k] = own[1 + k];
                if (k >= nmb) continue;
                const H2 h = own[1 + k], a = own[k], b = abv[k + 1], c = abv[k + 2];
                const bool dyn = (h.y & 0xFF00u) != 0u;
                bool uni;                                /* as decide_segment in the kernel */
                if (skip_spec) uni = ((h.x ^ a.x) | (h.y ^ a.y) | (h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
                else if (row == 0) uni = ((h.x ^ a.x) | (h.y ^ a.y)) == 0u;
                else uni = ((h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
                uni = uni && !dyn;
                u32 kd;
                if (uni && ((s + k) % 3)) kd = decide_mb_uniform(h.y, num_refs, skip_spec, T.code[k], T.len[k]);
                else if ((h.y & 0xFF00u) == 0) kd = decide_mb(h.x, h.y, a.x, a.y, b.x, b.y, c.x, c.y, num_refs, pred_spec, skip_spec, T.code[k], T.len[k]);
                else {
                    const u32 di = h.y >> 16;
                    T.len[k] = tab[di + 1] - tab[di] + (skip_spec ? 0u : 1u);
                    T.code[k] = tab[di];
                    kd = MBK_DYN;
                }
                T.kinds |= kd << (4 * k);
            }
        }
        /* prefixes */
        int prev = last_coded;
        for (int t = 0; t < n; t++) {
            th[t].prev0 = prev;
            if (skip_spec) prev = attach_runs(th[t].code, th[t].len, th[t].kinds, th[t].g0, prev);
        }
        const int last_coded_total = prev;
        /* trailer + scan */
        u32 tr_code = 1, tr_len = 1;
        if (last && skip_spec) {
            const u32 run_end = (u32)(W * Hh - 1 - last_coded_total);
            if (run_end > 0) { tr_code = ((run_end + 1) << 1) | 1; tr_len = 2u * flo32(run_end + 1u) + 2u; }
        }
        u32 run = bitpos;
        for (int t = 0; t < n; t++) {
            Th &T = th[t];
            T.Ls = 0;
            for (int k = 0; k < MPT; k++) T.Ls += T.len[k];
            T.cplx = kinds_complex(T.kinds);
            T.trailer = last && t == n - 1;
            T.L = T.Ls + (T.trailer ? tr_len : 0u);
            T.O = run; run += T.L;
        }
        const u32 end = run;
        const u32 total_bytes = (end + 7u) >> 3;
        const u32 w_end = end >> 5;
        const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;
        const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end;

        /* pre-flush */
        {
            const u32 wf_al = (bitpos >> 5) & ~3u;
            if (!tile0 && CH >= 8 && (CH & 3) == 0 && w_last - wbase >= CH && wf_al > wbase) {
                const u32 src = wf_al - wbase;
                emit_window(wbase << 2, (int)(src << 2));
                u32 c[4];
                for (int i = 0; i < 4; i++) c[i] = win[src + i];
                std::fill(win.begin(), win.end(), 0u);
                for (int i = 0; i < 4; i++) win[i] = c[i];
                wbase = wf_al;
            }
        }
        for (;;) {
            if (tile0)
                for (u32 i = 0; 32u * i < hdr_total; i++) {
                    u32 v = header_bits32(d.hdr_bits, d.nal_header, 32u * i);
                    const u32 left = hdr_total - 32u * i;
                    if (left < 32u) v &= ~(0xFFFFFFFFu >> left);
                    if (i - wbase < CH) win[i - wbase] |= v;
                }
            std::vector<Item> items;
            for (int t = 0; t < n; t++) {
                Th &T = th[t];
                if (!T.cplx && T.Ls) {
                    const int pick = (t + s0) % 3;
                    if (T.Ls <= 64 && pick == 0) simple_words_emu<2>(T.code, T.len, T.O, T.Ls, win.data(), wbase, CH);
                    else if (T.Ls <= 128 && pick <= 1) simple_words_emu<4>(T.code, T.len, T.O, T.Ls, win.data(), wbase, CH);
                    else simple_words_emu<8>(T.code, T.len, T.O, T.Ls, win.data(), wbase, CH);
                }
                u32 bit = T.O;
                int pv = T.prev0;
                if (T.cplx) {
                    for (int k = 0; k < MPT; k++) {
                        const u32 kd = (T.kinds >> (4 * k)) & 0xFu;
                        if (!T.len[k]) continue;
                        const u32 pre = skip_spec ? (u32)(T.g0 + k - pv) : 1u;
                        pv = T.g0 + k;
                        items.push_back(make_item(kd, T.code[k], T.len[k], bit, pre, T.own[k].x, T.own[k].y));
                        bit += T.len[k];
                    }
                } else bit += T.Ls;
                if (T.trailer) { Item it; it.x = bit | (IT_BITS << 30); it.y = 0; it.z = tr_code; it.w = tr_len; items.push_back(it); }
            }
            for (auto &it : items) {
                const u32 self[6] = {it.x, it.y, it.z, it.w, 0u, 0u};      /* the item as it lies in the queue */
                consume_item(it, self, win.data(), wbase, CH, num_refs, pay_words, n_pay_words);
            }

            const bool fits = w_last - wbase < CH;
            if (!(last || !fits)) break;
            const u32 e_hi = fits ? w_stop : wbase + CH;
            const u32 chunk_byte0 = wbase << 2;
            const int chunk_bytes = e_hi > wbase ? (int)((fits ? std::min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0;
            emit_window(chunk_byte0, chunk_bytes);
            std::fill(win.begin(), win.end(), 0u);
            if (fits) break;
            wbase += CH;
        }
        bitpos = end;
        last_coded = last_coded_total;
        if (last) size = total_bytes + ins_total;
        s0 += n;
    }
    return size;
}
