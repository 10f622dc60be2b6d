/*
 * coder_lean.cuh - the coder group of the reference-mode composer (skip OFF, reference predictor: the mode that is
 * bit-exact with the reference, and the one the benchmark measures).  Included by compose_kernels.cu.
 *
 * Same pipeline as the general coder (coder_main<true>, which keeps the SPEC modes), written for a small instruction
 * footprint: the kernel's speed follows the hit rate of the SM's 32 KB instruction cache, which four concurrent code
 * streams (two resident CTAs x coders / packers) overflow as soon as the hot loops are unrolled.  So:
 *
 *   decide   ONE rolled loop over the segment's 8 macroblocks.  A/B/D travel in registers from iteration to
 *            iteration, the own hint and C are read from the staged tile per iteration; there is one copy of the
 *            predictor in the code instead of eight.  Short codes go straight into a right-aligned 256-bit
 *            accumulator (2, 4 or 8 words are shifted, picked per warp from the longest string so far): no per-
 *            macroblock code / length arrays, no concatenation pass.
 *   splice   a thread's spliced macroblocks normally form one run with consecutive payload indices (a row of a
 *            dynamic rectangle): the thread keeps (first MB, count, first index) and pushes the run's items from the
 *            staged offset table alone.  The bits in front of and behind the run (<= 64) are OR-ed into the window
 *            directly.  Everything irregular - several runs, scattered indices, codes longer than 32 bits - takes
 *            the out-of-line path, which recomputes the thread's macroblocks from the dense map in global memory and
 *            turns every one of them into a queue item (the general coder's way).
 *   trailer  the slice's stop bit rides in the accumulator of the thread that owns the last macroblock.
 */
#ifndef H264SCROLL_CODER_LEAN_CUH
#define H264SCROLL_CODER_LEAN_CUH

/* staged hint `half` of 16-byte chunk c of the tile (c = 4 * slot + pair) */
__device__ __forceinline__ uint2 ld_hint(const uint4 *buf, u32 c, u32 half) {
    return ((const uint2 *)buf)[2u * hint_swizzle(c) + half];
}

/* a <<= n (0..32) over the low NWa words, then OR the code in */
template <int NWa>
__device__ __forceinline__ void acc_append(u32 (&a)[8], u32 code, u32 n) {
#pragma unroll
    for (int i = 8 - NWa; i < 7; i++) a[i] = fshl_c(a[i], a[i + 1], n);
    a[7] = fshl_c(a[7], 0u, n) | code;
}

/*
 * Bits [O, O + Ls) of the slice from the low NWc words of the accumulator.  Stores the whole words the thread owns;
 * hands back the partial head word H (word O >> 5, when O is not word aligned) and the last word T (word
 * (O + Ls - 1) >> 5) unless that is the head word.
 */
template <int NWc>
__device__ __forceinline__ void place_words(const u32 (&a)[8], u32 O, u32 Ls, u32 *win, u32 wbase, u32 &H, u32 &T, bool &has_t) {
    const u32 E = O + Ls;
    const u32 j0 = O >> 5, s = O & 31u, jl = (E - 1u) >> 5;
    const int r = (int)((32u - (E & 31u)) & 31u);
    u32 b[NWc + 1];
    b[0] = fshl(0u, a[8 - NWc], r);
#pragma unroll
    for (int i = 1; i < NWc; i++) b[i] = fshl(a[8 - NWc + i - 1], a[8 - NWc + i], r);
    b[NWc] = a[7] << r;
    const u32 lo_idx = j0 + (s ? 1u : 0u);                   /* first word this thread stores itself */
    H = 0u;
#pragma unroll
    for (int i = 0; i < NWc; i++) {
        const u32 idx = jl - (u32)(NWc - i);
        if ((int)(idx - lo_idx) >= 0 && idx - wbase < (u32)CH) win[idx - wbase] = b[i];
        if (idx == j0) H = b[i];
    }
    has_t = !(jl == j0 && s);
    T = b[NWc];
    if (!has_t) H = b[NWc];
    if (!s) H = 0u;
}

/*
 * Out of line: every macroblock of an irregular thread as a queue item, recomputed from the dense map in global memory
 * (neighbours by position; C falls back to D at the row's end, src/h264_writer.c:391-399).  Items whose tile index lies
 * in [r_lo, r_hi) are stored at queue[index - r_lo]; idx0 = tile index of the thread's first item, bit = its first bit.
 */
__device__ __noinline__ void irregular_items(const uint2 *__restrict__ fh, int W, int row, int x0, int nmb, int num_refs,
                                             const u32 *tab, u32 bit, u32 idx0, u32 r_lo, u32 r_hi, uint4 *queue) {
    const uint2 un = make_uint2(0u, HINT_Y_UNAVAIL);
    u32 idx = idx0;
#pragma unroll 1
    for (int k = 0; k < nmb; k++) {
        const int x = x0 + k;
        const long long g = (long long)row * W + x;
        const uint2 h = __ldg(fh + g);
        u32 code = 0u, len = 0u, kd;
        if (h.y & 0xFF00u) {
            const u32 di = h.y >> 16, b0 = tab[di];
            len = tab[di + 1] - b0 + 1u;
            code = b0;
            kd = MBK_DYN;
        } else {
            const uint2 a = x > 0 ? __ldg(fh + g - 1) : un;
            uint2 b = un, c = un;
            if (row > 0) {
                b = __ldg(fh + g - W);
                c = x + 1 < W ? __ldg(fh + g - W + 1) : (x > 0 ? __ldg(fh + g - W - 1) : un);
            }
            kd = decide_mb(h.x, h.y, a.x, a.y, b.x, b.y, c.x, c.y, num_refs, 0, 0, code, len);
        }
        if (idx - r_lo < r_hi - r_lo) {
            const Item it = make_item(kd, code, len, bit, 1u, h.x, h.y);
            queue[idx - r_lo] = make_uint4(it.x, it.y, it.z, it.w);
        }
        idx++;
        bit += len;
    }
}

__device__ __forceinline__ void coder_lean(ComposeSmem &sm, int f, const H264GpuFrameDesc *__restrict__ descs, int n_frames,
                                           const H264GpuMbHint *__restrict__ hints,
                                           const H264GpuDynDesc *__restrict__ dyns,
                                           const u32 *__restrict__ dyn_tab,
                                           const uint8_t *__restrict__ payload,
                                           uint8_t *__restrict__ out, u32 *__restrict__ sizes) {
    const int tid = threadIdx.x, lane = tid & 31, gw = tid >> 5;
    u32 spar = 0;                      /* which sm.scan buffer the next group scan uses */
    u32 ipar = 0;                      /* which sm.tile_items counter the current tile uses */
    Ring ring; ring.slot = 0; ring.confirmed = -1;   /* (warp-specialised build only) */
    PackState pst;                     /* (fused build only: this thread is a packer too) */
    pack_state_init(pst);
    u32 buf = 0;                       /* the window being filled */
    u32 items_used = 0;                /* items queued in the slot being filled */
    u32 tab_par = 0;                   /* phase of sm.tab_full the next staged table completes */

    FrameGeo geo = load_geo(descs, f, hints);
    stage_hints(sm.hbuf, geo, 0, min(geo.seg_max, geo.n_seg));
    cp_async_commit();

    while (f < n_frames) {
        /* ---- frame setup: the descriptor is read by every thread (one broadcast line) ------ */
        const uint4 *dq = (const uint4 *)(descs + f);
        const uint4 d3 = __ldg(dq + 3);
        const u32 dyn_id = d3.y;
        const int W = geo.W;
        const u32 hdr_total = 40u + (d3.w & 0xFFu);
        const int num_refs = (int)((d3.w >> 8) & 0xFFu);
        const u32 *tab = nullptr;          /* the frame's payload bit-offset table as decide reads it */
        bool tab_staged = false;
        if (dyn_id != H264GPU_NO_DYN) {
            const uint4 *yq = (const uint4 *)(dyns + dyn_id);
            const uint4 y0 = __ldg(yq), y1 = __ldg(yq + 1);
            tab = dyn_tab + ((u64)y0.z | ((u64)y0.w << 32));
            /* the table (n_mbs + 1 words) goes to shared memory with one bulk copy of the 16-byte chunks that hold it;
               every coder is past its last read of the previous frame's table (they all arrived at that frame's last
               scan barrier, which this thread has left) */
            const uintptr_t a0 = (uintptr_t)tab & ~(uintptr_t)15;
            const u32 t_skip = (u32)(((uintptr_t)tab - a0) >> 2);
            const u32 t_bytes = (4u * (t_skip + y1.x + 1u) + 15u) & ~15u;
            tab_staged = y1.x < (u32)(TABCAP - 4);
            if (tab_staged) {
                if (tid == 0) {
                    mbar_arrive_expect_tx(&sm.tab_full, t_bytes);
                    bulk_g2s(sm.tab, (const void *)a0, t_bytes, &sm.tab_full);
                }
                tab = sm.tab + t_skip;
            }
            const uint8_t *pa = payload + ((u64)y0.x | ((u64)y0.y << 32));
            if (FUSED) {
                /* the payload goes to shared memory now; it lands while the frame's first tile is composed (nobody reads the
                   previous frame's payload any more: that frame ended with an emit) */
                stage_payload(sm, pst, sizes + f, (const u32 *)pa, (y1.y + 3u) >> 2, tid == 0);
            } else if (tid == 32 && y1.y) {
                /* the packers fetch the payload when they meet the frame: send it on its way to L2 now */
                const uintptr_t p0 = (uintptr_t)pa & ~(uintptr_t)15;
                const u64 p_bytes = (((uintptr_t)pa - p0) + (u64)y1.y + 15ull) & ~15ull;
                bulk_prefetch_l2((const void *)p0, (u32)(p_bytes < (1ull << 20) ? p_bytes : (1ull << 20)));
            }
        } else if (FUSED) stage_payload(sm, pst, sizes + f, nullptr, 0u, false);
        const int kc = (W - 1) & 7;                            /* position of a row's last MB inside its segment */

        /* frame state every thread tracks identically (all of it follows from group-wide totals) */
        u32 bitpos = hdr_total;        /* slice bits composed so far (incl. the 40-bit Annex-B prefix) */
        u32 wbase = 0;                 /* frame word held by window word 0; words below it have left for the packers */
        bool first_emit = true;

        const int f_next = next_own_frame<false>(descs, f + (int)gridDim.x, n_frames, (int)gridDim.x);
        /* the next frame's dense hint map is sent to L2 a whole frame ahead: the tile copies then find it there */
        if (tid == 64 && f_next < n_frames) {
            const uint4 *nq = (const uint4 *)(descs + f_next);
            const uint4 n2 = __ldg(nq + 2), n3 = __ldg(nq + 3);
            const u64 n_bytes = (8ull * (n3.z & 0xFFFFu) * (n3.z >> 16)) & ~15ull;
            if (n_bytes) bulk_prefetch_l2((const uint2 *)hints + ((u64)n2.x | ((u64)n2.y << 32)), (u32)(n_bytes < (1ull << 20) ? n_bytes : (1ull << 20)));
        }

        /* a window (+ its queue) goes to the packers (see coder_main) */
        auto hand_over = [&](u32 flags, u32 chunk_bytes, u32 clear_words, u32 total_bytes, u32 carry = 0u) {
            if (tid == 0) {
                SlotMeta &m = sm.meta[FUSED ? 0 : ring.slot & 1];
                const uint4 e2 = __ldg(dq + 2), e3 = __ldg(dq + 3);
                m.fout = out + ((u64)e2.z | ((u64)e2.w << 32)); m.cap = e3.x;
                m.pay_words = nullptr; m.n_pay_words = 0;
                if (e3.y != H264GPU_NO_DYN) {
                    const uint4 *yq = (const uint4 *)(dyns + e3.y);
                    const uint4 y0 = __ldg(yq), y1 = __ldg(yq + 1);
                    m.pay_words = (const u32 *)(payload + ((u64)y0.x | ((u64)y0.y << 32)));
                    m.n_pay_words = (y1.y + 3u) >> 2;
                }
                m.size_ptr = sizes + f;
                m.buf = buf; m.n_items = items_used; m.flags = flags;
                m.wbase = wbase; m.chunk_bytes = chunk_bytes; m.clear_words = clear_words; m.total_bytes = total_bytes;
                m.carry = carry; m.num_refs = num_refs;
            }
            if (FUSED) {                                            /* every thread turns packer for this window */
                __syncthreads();
                process_slot(sm, sm.meta[0], sm.items[0], tid, pst);
                __syncthreads();
            } else {
                mbar_arrive(&sm.full[ring.slot & 1]);
                ring.slot++;
            }
            items_used = 0;
        };

        for (int s0 = 0; s0 < geo.n_seg;) {
            const int n = min(geo.seg_max, geo.n_seg - s0);
            const bool last = (s0 + n >= geo.n_seg);
            const bool tile0 = (s0 == 0);
            const uint4 *hb = sm.hbuf;

            /* ---- stage: this tile's hints have landed ---------- */
            cp_async_wait_all();
            group_sync<BAR_CODER, NC>();
            if (tile0 && tab_staged) mbar_wait(&sm.tab_full, tab_par);

            /* ---- the thread's segment ------------------------ */
            const int s = s0 + tid;
            const bool active = tid < n;
            const int row = seg_row(geo, active ? s : s0), col = (active ? s : s0) - row * geo.tpr;
            const int nmb = active ? min(MPT, W - MPT * col) : 0;
            const bool first = col == 0, lastseg = col == geo.tpr - 1;
            const u32 c_own = 4u * (u32)(tid + geo.tpr + 1), c_abv = 4u * (u32)(tid + 1);   /* first chunks of the own segment / the one above */
            const bool owns_trailer = last && tid == n - 1;

            /* ---- decide + code: two macroblocks (one 16-byte chunk of the staged tile) per iteration ---------------- */
            u32 acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0u;
            u32 Ls = 0;                /* bits in the accumulator */
            u32 Ldyn = 0;              /* bits of the spliced macroblocks (with their mb_skip_run bit) */
            u32 Lpre = 0;              /* accumulator bits in front of the spliced run */
            u32 run = 0;               /* k_first | n_run << 8 | irregular << 16 */
            u32 di_first = 0;
            {
                const bool row0 = row == 0;
                /* one macroblock against A, B, C (D where the row ends, src/h264_writer.c:391-399): its code and length */
                auto one_mb = [&](int k, const uint2 h, const uint2 a, const uint2 b, uint2 c, const uint2 d, u32 &code, u32 &len) {
                    if (lastseg && k == kc) c = d;
                    const bool present = k < nmb;
                    const bool dyn = (h.y & 0xFF00u) != 0u;
                    /* top row: A alone decides (src/h264_writer.c:407-417); else B == C == own hint: two refs match and
                       min(max(a, h), h) is h whatever A is */
                    const bool uni = row0 ? ((h.x ^ a.x) | (h.y ^ a.y)) == 0u
                                          : ((h.x ^ b.x) | (h.y ^ b.y) | (h.x ^ c.x) | (h.y ^ c.y)) == 0u;
                    code = 0u; len = 0u;
                    if (__all_sync(FULL, uni || dyn || !present)) {
                        if (present && !dyn) decide_mb_uniform(h.y, num_refs, 0, code, len);
                    } else if (present && !dyn) {
                        if (decide_mb(h.x, h.y, a.x, a.y, b.x, b.y, c.x, c.y, num_refs, 0, 0, code, len) == MBK_LONG) {
                            run |= 1u << 16;                           /* a code longer than 32 bits: the out-of-line path */
                            Ldyn += len; code = 0u; len = 0u;          /* (its bits are counted, not accumulated) */
                        }
                    }
                    if (__any_sync(FULL, present && dyn)) {
                        if (present && dyn) {
                            const u32 di = h.y >> 16;
                            const u32 n_run = (run >> 8) & 0xFFu;
                            Ldyn += tab[di + 1] - tab[di] + 1u;        /* (generic loads: the table is staged in shared memory when it fits) */
                            if (n_run == 0u) { run |= (u32)k | (1u << 8); di_first = di; Lpre = Ls + len; }
                            else if ((u32)k == (run & 0xFFu) + n_run && di == di_first + n_run) run += 1u << 8;
                            else run |= 1u << 16;                      /* a second run, or scattered payload indices */
                        }
                    }
                };
                const uint2 un = make_uint2(0u, HINT_Y_UNAVAIL);
                uint2 a = un, d = un;
                uint4 pa = make_uint4(0u, HINT_Y_UNAVAIL, 0u, HINT_Y_UNAVAIL);      /* chunk j of the segment above */
                if (active) {
                    if (!first) { a = ld_hint(hb, c_own - 1u, 1u); d = ld_hint(hb, c_abv - 1u, 1u); }
                    pa = hb[hint_swizzle(c_abv)];
                }
#pragma unroll 1
                for (int j = 0; j < MPT / 2; j++) {
                    uint4 o = make_uint4(0u, HINT_Y_UNAVAIL, 0u, HINT_Y_UNAVAIL), na = o;
                    if (active) {
                        o = hb[hint_swizzle(c_own + (u32)j)];
                        na = hb[hint_swizzle(c_abv + (u32)j + 1u)];
                    }
                    const uint2 h0 = make_uint2(o.x, o.y), h1 = make_uint2(o.z, o.w);
                    const uint2 b0 = make_uint2(pa.x, pa.y), b1 = make_uint2(pa.z, pa.w);
                    u32 code0, len0, code1, len1;
                    one_mb(2 * j, h0, a, b0, b1, d, code0, len0);
                    const u32 Ls0 = Ls;
                    Ls += len0;                                        /* (Lpre of a run that starts at the second macroblock) */
                    one_mb(2 * j + 1, h1, h0, b1, make_uint2(na.x, na.y), b0, code1, len1);
                    Ls = Ls0;
                    const u32 lmax = __reduce_max_sync(FULL, Ls + len0 + len1);
                    if (lmax <= 64u) { acc_append<2>(acc, code0, len0); acc_append<2>(acc, code1, len1); }
                    else if (lmax <= 128u) { acc_append<4>(acc, code0, len0); acc_append<4>(acc, code1, len1); }
                    else { acc_append<8>(acc, code0, len0); acc_append<8>(acc, code1, len1); }
                    Ls += len0 + len1;
                    a = h1; d = b1; pa = na;
                }
            }
            bool irregular = (run >> 16) & 1u;
            const u32 n_run = (run >> 8) & 0xFFu;
            if (n_run && Ls > 64u) irregular = true;                    /* (the bits around a run are OR-ed in as one 64-bit value) */
            const bool cplx = n_run != 0u || irregular;
            /* the slice's rbsp_stop_one_bit: in the accumulator, or as an item behind a thread that works through items */
            u32 tr_len = 0u;
            if (owns_trailer) {
                if (cplx) tr_len = 1u;
                else {
                    const u32 lmax = Ls + 1u;
                    if (lmax <= 64u) acc_append<2>(acc, 1u, 1u);
                    else if (lmax <= 128u) acc_append<4>(acc, 1u, 1u);
                    else acc_append<8>(acc, 1u, 1u);
                    Ls += 1u;
                }
            }

            /* ---- bit count of the thread, queue slots ---------------------------- */
            const u32 n_it = (irregular ? (u32)nmb : n_run) + (tr_len ? 1u : 0u);
            u32 ibase = 0;
            if (n_it) ibase = atomicAdd(&sm.tile_items[ipar], n_it);
            const u32 L = Ls + Ldyn + tr_len;

            /* ---- bit offsets: prefix sum of the threads' bit counts ----------------- */
            u32 total;
            const u32 O = bitpos + group_excl_scan<NWC, BAR_CODER>(L, sm.scan[spar], gw, total);
            spar ^= 1u;
            const u32 total_items = sm.tile_items[ipar];
            /* ---- stage: everybody is past its loads from the hint tile: start the copy of the next one ---- */
            FrameGeo geo_next = geo;
            if (!last) stage_hints(sm.hbuf, geo, s0 + n, min(geo.seg_max, geo.n_seg - s0 - n));
            else if (f_next < n_frames) {
                geo_next = load_geo(descs, f_next, hints);
                stage_hints(sm.hbuf, geo_next, 0, min(geo_next.seg_max, geo_next.n_seg));
            }
            cp_async_commit();
            if (tid == 0) sm.tile_items[ipar ^ 1u] = 0;            /* the next tile's counter; nobody adds to it before the next barrier */
            const u32 E = O + L;
            const u32 end = bitpos + total;
            const u32 total_bytes = (end + 7u) >> 3;                       /* meaningful when last */
            const u32 w_end = end >> 5;
            const u32 w_stop = last ? ((total_bytes + 3u) >> 2) : w_end;   /* frame words [.., w_stop) are complete after this tile */
            const u32 w_last = last ? (w_stop ? w_stop - 1u : 0u) : w_end; /* last word this tile writes to */
            const u32 warp_end = __shfl_sync(FULL, E, 31);

            /* ---- the tile does not fit behind what the window (or its queue) already holds: send that off first ---- */
            if (!tile0 && (w_last - wbase >= (u32)CH || items_used + total_items > (u32)LCAP)) {
                const u32 wf_al = (bitpos >> 5) & ~3u;             /* 16-byte group of the tile's first word */
                if (wf_al > wbase) {
                    const u32 src = wf_al - wbase;
                    if (!FUSED) ring_wait(sm, ring, ring.slot - 1);   /* the other window must be free: the packers move the last words over */
                    hand_over(SLOT_EMIT | (first_emit ? SLOT_FIRST : 0u), src << 2, src + 4u, 0u, src);
                    first_emit = false;
                    if (!FUSED) buf ^= 1u;
                    wbase = wf_al;
                } else if (items_used) {
                    hand_over(0u, 0u, 0u, 0u);                     /* nothing whole to emit yet: splice only, and wait for it */
                    if (!FUSED) ring_wait(sm, ring, ring.slot - 1);
                }
            }

            for (;;) {
                if (!FUSED) ring_wait(sm, ring, ring.slot - 2);    /* the window and the slot's queue are ours */
                u32 *win = sm.win[buf];
                /* ---- header: start code, NAL header byte, slice header bits ---------- */
                if (tile0 && tid < 10 && 32u * (u32)tid < hdr_total) {
                    const uint4 h0 = __ldg(dq), h1 = __ldg(dq + 1);
                    const u64 hdr[4] = {(u64)h0.x | ((u64)h0.y << 32), (u64)h0.z | ((u64)h0.w << 32),
                                        (u64)h1.x | ((u64)h1.y << 32), (u64)h1.z | ((u64)h1.w << 32)};
                    u32 v = header_bits32(hdr, d3.w >> 24, 32u * (u32)tid);
                    const u32 left = hdr_total - 32u * (u32)tid;
                    if (left < 32u) v &= ~(FULL >> left);
                    if ((u32)tid - wbase < (u32)CH && v) atomicOr(&win[(u32)tid - wbase], v);
                }

                /* ---- write: the accumulator + merge of the partial words between threads ---------- */
                {
                    const u32 Lc = cplx ? 0u : Ls;                         /* bits through the accumulator path */
                    const u32 Ec = O + Lc;
                    u32 H = 0u, T = 0u;
                    bool has_t = false;
                    const u32 lmax = __reduce_max_sync(FULL, Lc);
                    if (lmax <= 64u) { if (Lc) place_words<2>(acc, O, Lc, win, wbase, H, T, has_t); }
                    else if (lmax <= 128u) { if (Lc) place_words<4>(acc, O, Lc, win, wbase, H, T, has_t); }
                    else { if (Lc) place_words<8>(acc, O, Lc, win, wbase, H, T, has_t); }
                    const u32 key = O >> 5;
                    u32 tj = (Ec - 1u) >> 5;
                    if (!Lc) {
                        /* nothing through this path: the thread still closes the word its range ends in for the heads behind it */
                        tj = E >> 5;
                        has_t = L != 0u && (E & 31u) != 0u && (tj > key || (O & 31u) == 0u);
                    }
                    const bool plain = Lc != 0u && (has_t || (O & 31u) == 0u) && ((O & 31u) == 0u || tj > key);
                    if (__all_sync(FULL, plain)) {
                        /* every string reaches past its head word: a word is shared by two neighbours at most */
                        u32 v = T | __shfl_down_sync(FULL, H, 1);
                        if (lane == 31) v = T;
                        if (tj - wbase < (u32)CH) {
                            if (lane == 31) atomicOr(&win[tj - wbase], v);   /* the next warp's (or tile's) first head lands here */
                            else win[tj - wbase] = v;
                        }
                        if (lane == 0 && H != 0u && key - wbase < (u32)CH) atomicOr(&win[key - wbase], H);
                    } else {
                        /* general case: segmented suffix OR of the heads that share a word */
                        u32 X = H;
#pragma unroll
                        for (int dd = 1; dd < 32; dd <<= 1) {
                            const u32 v = __shfl_down_sync(FULL, X, dd);
                            const u32 kk = __shfl_down_sync(FULL, key, dd);
                            if (lane + dd < 32 && kk == key) X |= v;
                        }
                        const u32 Xn = __shfl_down_sync(FULL, X, 1);
                        const u32 kn = __shfl_down_sync(FULL, key, 1);
                        if (has_t && tj - wbase < (u32)CH) {
                            u32 v = T;
                            if (lane < 31 && kn == tj) v |= Xn;
                            if ((warp_end >> 5) == tj) atomicOr(&win[tj - wbase], v);   /* later warps (or the next tile) add to this word */
                            else win[tj - wbase] = v;
                        }
                        if (lane == 0 && X != 0u && key - wbase < (u32)CH) atomicOr(&win[key - wbase], X);
                    }
                    /* a thread with one spliced run: its (<= 64) bits in front of and behind the run are OR-ed in */
                    if (n_run && !irregular && Ls) {
                        const u64 v = ((u64)acc[6] << 32) | acc[7];
                        const u32 Lpost = Ls - Lpre;
                        if (Lpre) win_or64(win, wbase, (u32)CH, O, v >> Lpost, Lpre);
                        if (Lpost) win_or64(win, wbase, (u32)CH, O + Lpre + Ldyn, Lpost < 64u ? v & ((1ull << Lpost) - 1ull) : v, Lpost);
                    }
                }

                /* ---- push: this tile's items into the slot's queue; a tile with more items than the queue has room for
                        goes in rounds, each sent ahead as a splice-only hand-over ----------- */
                for (u32 r_lo = 0;;) {
                    const u32 r_hi = min(total_items, r_lo + ((u32)LCAP - items_used));
                    if (n_it && ibase < r_hi && ibase + n_it > r_lo) {
                        uint4 *queue = sm.items[FUSED ? 0 : ring.slot & 1] + items_used;
                        u32 idx = ibase, bit = O;
                        if (irregular) {
                            irregular_items(geo.fh, W, row, MPT * col, nmb, num_refs, tab, O, ibase, r_lo, r_hi, queue);
                            idx += (u32)nmb;
                            bit += Ls + Ldyn;
                        } else {
                            /* the run's items from the offset table: '1' (mb_skip_run = 0) + the macroblock's payload bits */
                            bit += Lpre;
                            u32 b0 = tab[di_first];
#pragma unroll 1
                            for (u32 i = 0; i < n_run; i++) {
                                const u32 b1 = tab[di_first + i + 1u];
                                if (idx - r_lo < r_hi - r_lo)
                                    queue[idx - r_lo] = make_uint4(bit | (IT_DYN << 30), 1u, b0, b1 - b0);
                                idx++;
                                bit += b1 - b0 + 1u;
                                b0 = b1;
                            }
                            bit += Ls - Lpre;
                        }
                        if (tr_len && idx - r_lo < r_hi - r_lo)
                            queue[idx - r_lo] = make_uint4(bit | (IT_BITS << 30), 0u, 1u, 1u);
                    }
                    items_used += r_hi - r_lo;
                    if (r_hi >= total_items) break;
                    hand_over(0u, 0u, 0u, 0u);                             /* (the same window goes on) */
                    if (!FUSED) ring_wait(sm, ring, ring.slot - 1);
                    r_lo = r_hi;
                }

                const bool fits = w_last - wbase < (u32)CH;                /* the tile's last word lies inside the window */
                if (!(last || !fits)) break;                               /* keep filling the window with the next tile */

                /* ---- the window is full or the frame ends: frame words [wbase, e_hi) go to the packers --------- */
                const u32 e_hi = fits ? w_stop : wbase + CH;
                const u32 chunk_byte0 = wbase << 2;
                const u32 chunk_bytes = e_hi > wbase ? ((fits ? min(e_hi << 2, total_bytes) : (e_hi << 2)) - chunk_byte0) : 0u;
                hand_over(SLOT_EMIT | (first_emit ? SLOT_FIRST : 0u) | (fits ? SLOT_LAST : 0u), chunk_bytes,
                          min((u32)CH, w_last - wbase + 1u), total_bytes);
                first_emit = false;
                if (!FUSED) buf ^= 1u;
                if (fits) break;                                           /* (frame done) */
                wbase += CH;                                               /* a full window has left; the tile goes on in the next one */
            }

            bitpos = end;
            ipar ^= 1u;
            s0 += n;
            if (last) { geo = geo_next; break; }
        }
        if (tab_staged) tab_par ^= 1u;
        f = f_next;
    }
    cp_async_wait_all();
    if (!FUSED) {                                                          /* tell the packers to leave */
        ring_wait(sm, ring, ring.slot - 2);
        if (tid == 0) sm.meta[ring.slot & 1].flags = SLOT_EXIT;
        mbar_arrive(&sm.full[ring.slot & 1]);
    }
}

#endif
