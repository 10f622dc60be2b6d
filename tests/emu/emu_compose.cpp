/*
 * emu_compose.cpp - TEST ONLY.  Sequential CPU emulation of the tile pipeline of
 * csrc/compose_kernels.cu (load -> decide -> skiprun -> scan -> write -> splice -> emit),
 * built on the very same mb_codec.cuh helpers the kernel uses (decide_inter, attach_prefix,
 * BitPut, epb_group_mask, realign_selector ...).  "Threads" are walked one after the other;
 * the warp-shuffle merge of partial words becomes a plain OR into the window.  It lets the
 * CPU test-suite check the bit-level logic against the oracle without a GPU, with tiny
 * thread counts / window sizes that force every boundary case.
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

static u32 pay32(const u32 *pw, u32 n_words, u32 p) { return payload_bits32(pw, n_words, p); }

static u32 byte_perm(u32 a, u32 b, u32 sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    u32 r = 0;
    for (int k = 0; k < 4; k++) r |= (u32)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
}

extern "C" uint32_t emu_compose_frame(const H264GpuFrameDesc *desc, const H264GpuMbHint *hints_base,
                                      const H264GpuDynDesc *dyns, const uint32_t *dyn_tab,
                                      const uint8_t *payload, uint8_t *out_base,
                                      int NTHREADS, int CH) {
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
    const int TILE = NTHREADS * MPT;
    const u32 hdr_total = 40u + d.hdr_nbits;

    u32 bitpos = hdr_total, carry = 0, ins_total = 0, zrun = 0;
    int last_coded = -1;
    u32 size = 0;
    const H2 zero_h = {0, 0};
    auto H = [&](int g) -> H2 { return (g >= 0 && g < N) ? fh[g] : zero_h; };

    std::vector<u64> code(TILE);
    std::vector<u32> meta(TILE), dynlen(NTHREADS), Lv(NTHREADS), Ov(NTHREADS);
    std::vector<int> previn(NTHREADS);
    std::vector<u32> win(CH + 8);

    for (int t0 = 0; t0 < N; t0 += TILE) {
        const bool last = (t0 + TILE >= N), tile0 = (t0 == 0);
        /* decide */
        for (int t = 0; t < NTHREADS; t++) {
            const int g0 = t0 + t * MPT;
            const int nmb = std::min(std::max(N - g0, 0), MPT);
            dynlen[t] = 0;
            for (int k = 0; k < MPT; k++) {
                u64 c = 0; u32 m = 0;
                if (k < nmb) {
                    const int g = g0 + k, y = g / W, x = g - y * W;
                    const H2 h = H(g);
                    if (((h.y >> 8) & 0xFF) == H264GPU_MB_INTER) {
                        const int avA = x > 0, avB = y > 0;
                        const bool hasC = x + 1 < W;
                        const int avC = avB && (hasC || x > 0);
                        const H2 a = H(g - 1), b = H(g - W), cc = hasC ? H(g - W + 1) : H(g - W - 1);
                        m = decide_inter(h.x, h.y, a.x, a.y, avA, b.x, b.y, avB, cc.x, cc.y, avC,
                                         num_refs, pred_spec, skip_spec, c);
                    } else {
                        const u32 di = h.y >> 16;
                        dynlen[t] += tab[di + 1] - tab[di];
                        m = META_DYN | META_CODED;
                        c = (u64)di << 48;
                    }
                }
                code[t * MPT + k] = c; meta[t * MPT + k] = m;
            }
        }
        /* prefixes */
        int prev = last_coded;
        for (int t = 0; t < NTHREADS; t++) {
            previn[t] = prev;
            const int g0 = t0 + t * MPT;
            for (int k = 0; k < MPT; k++) {
                u32 &m = meta[t * MPT + k];
                if (!(m & META_CODED)) continue;
                const u32 pre = skip_spec ? (u32)(g0 + k - prev) : 1u;
                attach_prefix(code[t * MPT + k], m, pre, 63 - 2 * clz32(pre));
                prev = g0 + k;
            }
        }
        const int last_coded_total = prev;
        /* trailer + scan */
        u32 tr_code = 1, tr_len = 1;
        if (last && skip_spec) {
            const u32 run_end = (u32)(N - 1 - last_coded_total);
            if (run_end > 0) { const int pn = 63 - 2 * clz32(run_end + 1); tr_code = ((run_end + 1) << 1) | 1; tr_len = pn + 1; }
        }
        u32 run = bitpos;
        int trailer_thread = -1;
        for (int t = 0; t < NTHREADS; t++) {
            const int g0 = t0 + t * MPT;
            const int nmb = std::min(std::max(N - g0, 0), MPT);
            const bool owns = last && nmb > 0 && g0 + nmb == N;
            if (owns) trailer_thread = t;
            u32 L = dynlen[t] + (owns ? tr_len : 0);
            for (int k = 0; k < MPT; k++) L += meta[t * MPT + k] & META_LEN_MASK;
            Lv[t] = L; Ov[t] = run; run += L;
        }
        const u32 end = run;
        const u32 w_lo = bitpos >> 5;
        const u32 total_bytes = (end + 7u) >> 3;
        const u32 w_end = end >> 5;
        const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;
        const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end;

        bool first_pass = true;
        u32 wbase_last = 0;
        for (u32 wbase = tile0 ? 0u : w_lo; wbase <= w_last; wbase += CH) {
            std::fill(win.begin(), win.end(), 0u);
            if (first_pass && !tile0) win[0] = carry;
            first_pass = false;
            wbase_last = wbase;
            struct Copy { u32 j, src, n; };
            std::vector<Copy> copies;
            for (int t = 0; t < NTHREADS; t++) {
                const int g0 = t0 + t * MPT;
                BitPut bp;
                bp_init(bp, (tile0 && t == 0) ? 0u : Ov[t]);
                const u32 key = bp.j;
                if (tile0 && t == 0)
                    for (u32 p = 0; p < hdr_total; p += 32) {
                        const u32 n = std::min(32u, hdr_total - p);
                        bp_put(bp, (u64)(header_bits32(d.hdr_bits, d.nal_header, p) >> (32 - n)), n, win.data(), wbase, CH);
                    }
                int pv = previn[t];
                for (int k = 0; k < MPT; k++) {
                    const u32 m = meta[t * MPT + k];
                    const u64 c = code[t * MPT + k];
                    if (!(m & META_CODED)) continue;
                    const u32 n = m & META_LEN_MASK;
                    if (!(m & (META_DYN | META_LONG | META_SPLIT))) bp_put(bp, c, n, win.data(), wbase, CH);
                    else {
                        const u32 pre = skip_spec ? (u32)(g0 + k - pv) : 1u;
                        const u32 pn = 63 - 2 * clz32(pre);
                        bp_put(bp, pre, pn, win.data(), wbase, CH);
                        if (m & META_SPLIT) bp_put(bp, c, n - pn, win.data(), wbase, CH);
                        else if (m & META_LONG) {
                            MbEntry e;
                            int bl = encode_p16x16_body((int)(m >> META_REF_SHIFT), long_dx(c), long_dy(c), num_refs, e);
                            bp_put(bp, e.w1, 32, win.data(), wbase, CH);
                            bp_put(bp, e.w2, 32, win.data(), wbase, CH);
                            bl -= 64;
                            bp_put(bp, e.w3 >> (32 - bl), bl, win.data(), wbase, CH);
                        } else {
                            const u32 di = (u32)(c >> 48);
                            const u32 b0 = tab[di], len = tab[di + 1] - b0;
                            const u32 hb = std::min(len, (32u - bp.fill) & 31u);
                            if (hb) bp_put(bp, pay32(pay_words, n_pay_words, b0) >> (32 - hb), hb, win.data(), wbase, CH);
                            const u32 rest = len - hb, nw = rest >> 5, tb = rest & 31;
                            if (nw) { copies.push_back({bp.j, b0 + hb, nw}); bp.j += nw; }
                            if (tb) bp_put(bp, pay32(pay_words, n_pay_words, b0 + len - tb) >> (32 - tb), tb, win.data(), wbase, CH);
                        }
                    }
                    pv = g0 + k;
                }
                if (t == trailer_thread) bp_put(bp, tr_code, tr_len, win.data(), wbase, CH);
                if (bp.fill) {
                    if (bp.head) bp.H = bp.acc;
                    else if (bp.j - wbase < (u32)CH) win[bp.j - wbase] |= bp.acc;
                }
                if (bp.H && key - wbase < (u32)CH) win[key - wbase] |= bp.H;
            }
            for (auto &c : copies)
                for (u32 i = 0; i < c.n; i++)
                    if (c.j + i - wbase < (u32)CH) win[c.j + i - wbase] = pay32(pay_words, n_pay_words, c.src + 32 * i);

            /* emit */
            const u32 e_hi = std::min(wbase + (u32)CH, w_stop);
            const u32 chunk_byte0 = wbase << 2;
            const int chunk_bytes = e_hi > wbase ? (int)((last ? std::min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0;
            for (int q0 = 0; q0 * 16 < chunk_bytes; q0 += NTHREADS) {
                std::vector<u32> tr(NTHREADS), zs(NTHREADS), insv(NTHREADS), rins(NTHREADS);
                std::vector<int> vbs(NTHREADS);
                for (int t = 0; t < NTHREADS; t++) {
                    const int q = q0 + t;
                    const int vb = std::min(std::max(chunk_bytes - 16 * q, 0), 16);
                    u32 w[4] = {~0u, ~0u, ~0u, ~0u};
                    if (vb > 0) for (int i = 0; i < 4; i++) w[i] = win[4 * q + i];
                    u32 z = zero_nib(w[0]) | (zero_nib(w[1]) << 4) | (zero_nib(w[2]) << 8) | (zero_nib(w[3]) << 12);
                    const u32 valid = (1u << vb) - 1u;
                    z &= valid;
                    if (chunk_byte0 == 0 && q == 0) z &= ~0x1Fu;
                    zs[t] = z; vbs[t] = vb;
                    tr[t] = epb_group_trail(z, std::max(vb, 1), 0) | (z == valid ? 0x100u : 0u);
                }
                u32 pass_ins = 0, zrun_next = zrun;
                for (int t = 0; t < NTHREADS; t++) {
                    const int q = q0 + t, vb = vbs[t];
                    if (vb == 0) continue;
                    u32 w[5];
                    for (int i = 0; i < 5; i++) w[i] = win[4 * q + i];
                    const u32 valid = (1u << vb) - 1u, z = zs[t];
                    u32 keep = valid;
                    if (chunk_byte0 == 0 && q == 0) keep &= ~0x1Fu;
                    const bool b0_le3 = (w[0] >> 26) == 0 && (keep & 1);
                    const bool need = b0_le3 || (z & 1);
                    u32 rin = 0;
                    if (need) {
                        int u = t - 1;
                        for (;;) {
                            if (u < 0) { rin += zrun; break; }
                            rin += tr[u] & 0xFF;
                            if (!(tr[u] & 0x100)) break;
                            u--;
                        }
                    }
                    u32 ins = 0;
                    if (need || (z & (z << 1))) {
                        u32 m3 = zero_nib(w[0] & 0xFCFCFCFCu) | (zero_nib(w[1] & 0xFCFCFCFCu) << 4) |
                                 (zero_nib(w[2] & 0xFCFCFCFCu) << 8) | (zero_nib(w[3] & 0xFCFCFCFCu) << 12);
                        ins = epb_group_mask(z, m3 & keep, vb, rin);
                    }
                    insv[t] = ins; rins[t] = rin;
                    const int vb_next = t == NTHREADS - 1 ? 0 : std::min(std::max(chunk_bytes - 16 * (q + 1), 0), 16);
                    if (vb_next == 0) zrun_next = epb_group_trail(z, vb, rin);
                }
                u32 before = ins_total;
                for (int t = 0; t < NTHREADS; t++) {
                    const int q = q0 + t, vb = vbs[t];
                    if (vb == 0) continue;
                    const u32 cnt = (u32)__builtin_popcount(insv[t]);
                    const u32 D = chunk_byte0 + 16u * q + before;
                    auto is_clean = [&](int u) { return u >= 0 && u < NTHREADS && vbs[u] == 16 && insv[u] == 0; };
                    const bool clean = is_clean(t), prev_clean = is_clean(t - 1), next_clean = is_clean(t + 1);
                    u32 w[5];
                    for (int i = 0; i < 5; i++) w[i] = win[4 * q + i];
                    if (clean) {
                        const u32 s = (4u - (D & 3u)) & 3u;
                        if (s && !prev_clean)
                            for (u32 k = 0; k < s; k++) if (D + k < cap) fout[D + k] = (uint8_t)(w[0] >> (24 - 8 * k));
                        const u32 sel = realign_selector(s);
                        u32 o[4];
                        for (int i = 0; i < 4; i++) o[i] = byte_perm(w[i], w[i + 1], sel);
                        const u32 nfull = (s == 0 || next_clean) ? 4 : 3;
                        for (u32 i = 0; i < nfull; i++)
                            if (D + s + 4 * i + 4 <= cap) memcpy(fout + D + s + 4 * i, &o[i], 4);
                            else for (u32 k = 0; k < 4; k++) if (D + s + 4 * i + k < cap) fout[D + s + 4 * i + k] = (uint8_t)(o[i] >> (8 * k));
                        if (nfull == 3)
                            for (u32 k = 0; k < 4u - s; k++) if (D + s + 12 + k < cap) fout[D + s + 12 + k] = (uint8_t)(o[3] >> (8 * k));
                    } else {
                        u32 pos = D;
                        for (int k = 0; k < vb; k++) {
                            if ((insv[t] >> k) & 1) { if (pos < cap) fout[pos] = 3; pos++; }
                            if (pos < cap) fout[pos] = (uint8_t)(w[k >> 2] >> (24 - 8 * (k & 3)));
                            pos++;
                        }
                    }
                    before += cnt; pass_ins += cnt;
                }
                ins_total += pass_ins;
                zrun = zrun_next;
            }
        }
        bitpos = end;
        if (!last) carry = win[w_end - wbase_last];
        last_coded = last_coded_total;
        if (last) size = total_bytes + ins_total;
    }
    return size;
}
