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
    const int uniform_mode = 1;      /* take the kernel's uniform-neighbourhood shortcut where it applies */
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
                        const u32 hkey = h.y & 0xFFFFu;
                        const bool uni = avA && avC && a.x == h.x && b.x == h.x && cc.x == h.x &&
                                         (a.y & 0xFFFFu) == hkey && (b.y & 0xFFFFu) == hkey && (cc.y & 0xFFFFu) == hkey;
                        if (uni && (uniform_mode & 1)) m = decide_uniform(h.y, num_refs, skip_spec, c);
                        else m = decide_inter(h.x, h.y, a.x, a.y, avA, b.x, b.y, avB, cc.x, cc.y, avC,
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
        /* merge runs of plain codes into <= 64-bit items; thread bit counts */
        for (int t = 0; t < NTHREADS; t++) {
            u64 acc = 0; u32 accn = 0, L = dynlen[t];
            u64 *cd = &code[t * MPT]; u32 *mt = &meta[t * MPT];
            for (int k = 0; k < MPT; k++) {
                const u64 c = cd[k]; const u32 m = mt[k];
                L += m & META_LEN_MASK;
                cd[k] = 0; mt[k] = 0;
                if (m & (META_DYN | META_LONG)) {
                    if (k > 0 && accn) { cd[k - 1] = acc; mt[k - 1] = accn; acc = 0; accn = 0; }
                    cd[k] = c; mt[k] = m;
                } else if (m & META_CODED) {
                    u64 sc = 0; u32 sn = 0;
                    const bool seal = merge_plain(acc, accn, c, m & META_LEN_MASK, sc, sn);
                    if (k > 0 && seal) { cd[k - 1] = sc; mt[k - 1] = sn; }
                }
            }
            if (accn) { cd[MPT - 1] = acc; mt[MPT - 1] = accn; }
            Lv[t] = L;
        }
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
            Lv[t] += owns ? tr_len : 0;
            Ov[t] = run; run += Lv[t];
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
                for (int k = 0; k < MPT; k++) {
                    const u32 m = meta[t * MPT + k];
                    const u64 c = code[t * MPT + k];
                    if (!m) continue;
                    if (!(m & (META_DYN | META_LONG))) bp_put(bp, c, m, win.data(), wbase, CH);
                    else if (m & META_LONG) {
                        const u32 pre = long_prefix(c);
                        bp_put(bp, pre, 63 - 2 * clz32(pre), win.data(), wbase, CH);
                        MbEntry e;
                        int bl = encode_p16x16_body((int)(m >> META_REF_SHIFT), long_dx(c), long_dy(c), num_refs, e);
                        const u32 ws[3] = {e.w1, e.w2, e.w3};
                        for (int i = 0; i < 3 && bl > 0; i++, bl -= 32) {
                            const u32 n = (u32)std::min(bl, 32);
                            bp_put(bp, ws[i] >> (32 - n), n, win.data(), wbase, CH);
                        }
                    } else {
                        const u32 pre = dyn_prefix(c);
                        bp_put(bp, pre, 63 - 2 * clz32(pre), win.data(), wbase, CH);
                        const u32 di = dyn_index(c);
                        const u32 b0 = tab[di], len = tab[di + 1] - b0;
                        const u32 hb = std::min(len, (32u - bp.fill) & 31u);
                        if (hb) bp_put(bp, pay32(pay_words, n_pay_words, b0) >> (32 - hb), hb, win.data(), wbase, CH);
                        const u32 rest = len - hb, nw = rest >> 5, tb = rest & 31;
                        if (nw) { copies.push_back({bp.j, b0 + hb, nw}); bp.j += nw; }
                        if (tb) bp_put(bp, pay32(pay_words, n_pay_words, b0 + len - tb) >> (32 - tb), tb, win.data(), wbase, CH);
                    }
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
        }
        bitpos = end;
        if (!last) carry = win[w_end - wbase_last];
        last_coded = last_coded_total;
        if (last) size = total_bytes + ins_total;
    }
    return size;
}
