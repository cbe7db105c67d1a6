// sweep.cuh -- member sweeps: MIN / MAX / MEDIAN / overlap bp (+ SUM / MEAN when the exact prefix
// path is unavailable) and the emit pass of left_overlaps.  Included by query.cu.
//
// k_tile<M_UBLB> has already resolved, per left,
//     ub = #{r.start < l.end}   pp = #{r.start < l.start}   lb = #{r.end <= l.start}
// so in start order the group is:  every right in [pp, ub)  (it starts inside the left, so it
// overlaps -- no need to look at its end)  plus  pp - lb "straddlers" somewhere below pp with
// end > l.start, found by walking backward over `ends`, skipping 64-blocks whose block-max-end
// (the replacement of coitrees' subtree_last) cannot overlap, until pp - lb of them were seen.
//
// Work assignment: a group of G lanes per left (G = 8 for the usual tens of members, 32 when a
// tile averages more than 48 members per left), one CTA per tile of 1024 lefts, so a warp keeps
// 32 / G lefts in flight.

constexpr int SW2_TPB = 256;
constexpr u32 MED2_CAP = 128;  // scores a group keeps in shared memory for the median (G = 8)
constexpr u32 MED32_CAP = 512; // ... (G = 32)

template <int G>
__device__ __forceinline__ u32 group_mask() {
    return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
}

template <int G, bool WIDTHS>  // WIDTHS: overlap bp / weighted mean are wanted (they need the members' starts)
__device__ __forceinline__ void sweep_tile(const SeqDev *__restrict__ sd, u64 g_lo, u64 g_hi, const u32 *__restrict__ ls,
                                           const u32 *__restrict__ le, const u32 *__restrict__ ubv, const u32 *__restrict__ ppv,
                                           const u32 *__restrict__ lbv, const SweepOut &o, double *s_buf /* [groups][CAP] */,
                                           double *s_med /* [groups][2] */) {
    constexpr u32 CAP = G == 32 ? MED32_CAP : MED2_CAP;
    constexpr int NGROUPS = SW2_TPB / G;
    const u32 glane = threadIdx.x & (G - 1), gidx = threadIdx.x / G;
    const u32 gmask = group_mask<G>();
    const u32 gshift = (threadIdx.x & 31) & ~(G - 1);
    const u32 lt = (1u << glane) - 1;
    const u32 *__restrict__ ends = sd->ends;
    const u32 *__restrict__ starts = sd->starts;
    const double *__restrict__ score = sd->score_a;
    const u32 *__restrict__ vbits = sd->vbits_a;
    const u32 *__restrict__ bme = sd->bme;
    const bool has_scores = sd->has_scores != 0, has_nulls = sd->has_nulls != 0;
    const bool sweep_sums = sd->sum_mode == GRB_SUM_SWEEP;
    u64 *kbuf = reinterpret_cast<u64 *>(s_buf) + (size_t)gidx * CAP;  // order-preserving integer images of the scores

    // Warp-synchronous walk (see sweep3.cuh): the groups of a warp take every step together, so each vote is one
    // full-warp instruction instead of one per group mask.
    const u64 g_warp = g_lo + (threadIdx.x >> 5) * (32 / G);
    const u32 rounds = g_warp < g_hi ? (u32)((g_hi - g_warp + NGROUPS - 1) / NGROUPS) : 0u;
    u64 g = g_lo + gidx;
    for (u32 round = 0; round < rounds; round++, g += NGROUPS) {
        const bool live = g < g_hi;
        u32 u = 0, p = 0, l = 0, lstart = 0, lend = 0;
        if (live) {
            u = ubv[g];
            p = ppv[g];
            l = lbv[g];
            lstart = ls[g];
            lend = le[g];
        }
        u32 nvalid = 0;
        double sum = 0.0, mn = 0.0, mx = 0.0;
        bool have = false;
        u64 bp = 0, wtot = 0;  // overlap bp of all members / of the members with a score
        double wsum = 0.0;     // sum of score * overlap bp
        // ---- members that start inside the left: [p, u), ascending
        const u32 fw = u > p ? u - p : 0u;
        const u32 fw_steps = __reduce_max_sync(GRB_FULL, (fw + G - 1) / G);
        for (u32 it = 0; it < fw_steps; it++) {
            const u32 j = p + it * G + glane;
            const bool in = j < u;
            bool v = in && has_scores;
            if (v && has_nulls) v = (vbits[j >> 5] >> (j & 31)) & 1u;
            const u32 vm = (__ballot_sync(GRB_FULL, v) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            u32 wid = 0;
            if (WIDTHS && in) {
                wid = min(lend, ends[j]) - starts[j];
                bp += wid;
            }
            if (v) {
                const double x = score[j];
                sum += x;
                if (WIDTHS) {
                    wsum += x * (double)wid;
                    wtot += wid;
                }
                if (isfinite(x)) {
                    mn = have ? fmin(mn, x) : x;
                    mx = have ? fmax(mx, x) : x;
                    have = true;
                }
                if (o.want_median) {
                    const u32 slot = nvalid + __popc(vm & lt);
                    if (slot < CAP) kbuf[slot] = f64_key(x);
                }
            }
            nvalid += __popc(vm);
        }
        // ---- straddlers: below p, end > l.start; exactly p - l of them
        const u32 need = p - l;
        u32 found = 0;
        long long cb = p ? (long long)((p - 1) & ~(u32)(G - 1)) : -1;
        while (true) {
            const bool act = found < need && cb >= 0;
            if (!__any_sync(GRB_FULL, act)) break;
            const bool look = act && bme[cb >> GRB_BME_SHIFT] > lstart;
            const u32 j = (u32)cb + glane;
            const bool in = look && j < p;
            const u32 e = in ? ends[j] : 0u;
            const bool hit = in && e > lstart;
            const u32 hm = (__ballot_sync(GRB_FULL, hit) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            bool v = hit && has_scores;
            if (v && has_nulls) v = (vbits[j >> 5] >> (j & 31)) & 1u;
            const u32 vm = (__ballot_sync(GRB_FULL, v) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            u32 wid = 0;
            if (WIDTHS && hit) {
                wid = min(lend, e) - lstart;
                bp += wid;
            }
            if (v) {
                const double x = score[j];
                sum += x;
                if (WIDTHS) {
                    wsum += x * (double)wid;
                    wtot += wid;
                }
                if (isfinite(x)) {
                    mn = have ? fmin(mn, x) : x;
                    mx = have ? fmax(mx, x) : x;
                    have = true;
                }
                if (o.want_median) {
                    const u32 slot = nvalid + __popc(vm & lt);
                    if (slot < CAP) kbuf[slot] = f64_key(x);
                }
            }
            found += __popc(hm);
            nvalid += __popc(vm);
            if (act) cb = look ? cb - G : ((cb >> GRB_BME_SHIFT) << GRB_BME_SHIFT) - G;
        }
        // ---- group reductions (fixed shuffle order: deterministic)
#pragma unroll
        for (int s = G / 2; s; s >>= 1) {
            sum += __shfl_xor_sync(GRB_FULL, sum, s);
            const double omn = __shfl_xor_sync(GRB_FULL, mn, s), omx = __shfl_xor_sync(GRB_FULL, mx, s);
            const bool ohave = __shfl_xor_sync(GRB_FULL, (int)have, s);
            if (ohave) {
                mn = have ? fmin(mn, omn) : omn;
                mx = have ? fmax(mx, omx) : omx;
                have = true;
            }
            if (WIDTHS) {
                bp += __shfl_xor_sync(GRB_FULL, bp, s);
                wsum += __shfl_xor_sync(GRB_FULL, wsum, s);
                wtot += __shfl_xor_sync(GRB_FULL, wtot, s);
            }
        }
        // ---- median of up to CAP values: quickselect inside the group.  Elements are ordered by (key, slot); a
        // round counts the elements below the pivot (one pass, each lane over its strided share) and, in the same
        // pass, remembers the first still-active element on either side as the next pivot.
        bool heavy = false;
        double med_lo = 0.0, med_hi = 0.0;
        if (o.want_median && nvalid) {
            if (nvalid > CAP) {
                heavy = true;
            } else {
                __syncwarp(gmask);
                const u32 m = nvalid, k1 = (m - 1) >> 1;
                // active range: (lo_k, lo_i) < element < (hi_k, hi_i), exclusive; starts unbounded
                u64 lo_k = 0, hi_k = ~0ull;
                u32 lo_i = 0, hi_i = 0xffffffffu;
                bool has_lo = false;
                u32 piv = 0;  // slot of the pivot
                u64 res_k = 0;
                u32 res_i = 0;
                while (true) {
                    const u64 pk = kbuf[piv];
                    u32 cnt = 0, next_below = 0xffffffffu, next_above = 0xffffffffu;
                    for (u32 e = glane; e < m; e += G) {
                        const u64 y = kbuf[e];
                        const bool below = y < pk || (y == pk && e < piv);
                        const bool above = y > pk || (y == pk && e > piv);
                        cnt += below ? 1u : 0u;
                        const bool gt_lo = !has_lo || y > lo_k || (y == lo_k && e > lo_i);
                        const bool lt_hi = y < hi_k || (y == hi_k && e < hi_i);
                        if (below && gt_lo) next_below = min(next_below, e);
                        if (above && lt_hi) next_above = min(next_above, e);
                    }
#pragma unroll
                    for (int s2 = G / 2; s2; s2 >>= 1) {
                        cnt += __shfl_xor_sync(gmask, cnt, s2);
                        next_below = min(next_below, __shfl_xor_sync(gmask, next_below, s2));
                        next_above = min(next_above, __shfl_xor_sync(gmask, next_above, s2));
                    }
                    if (cnt == k1) {
                        res_k = pk;
                        res_i = piv;
                        break;
                    }
                    if (cnt < k1) {
                        lo_k = pk;
                        lo_i = piv;
                        has_lo = true;
                        piv = next_above;
                    } else {
                        hi_k = pk;
                        hi_i = piv;
                        piv = next_below;
                    }
                }
                med_lo = key_f64(res_k);
                med_hi = med_lo;
                if (!(m & 1)) {
                    // even count: the upper middle is the smallest element above (res_k, res_i)
                    u64 bk = ~0ull;
                    u32 bi = 0xffffffffu;
                    for (u32 e = glane; e < m; e += G) {
                        const u64 y = kbuf[e];
                        const bool above = y > res_k || (y == res_k && e > res_i);
                        if (above && (y < bk || (y == bk && e < bi))) {
                            bk = y;
                            bi = e;
                        }
                    }
#pragma unroll
                    for (int s2 = G / 2; s2; s2 >>= 1) {
                        const u64 ok2 = __shfl_xor_sync(gmask, bk, s2);
                        const u32 oi2 = __shfl_xor_sync(gmask, bi, s2);
                        if (ok2 < bk || (ok2 == bk && oi2 < bi)) {
                            bk = ok2;
                            bi = oi2;
                        }
                    }
                    med_hi = key_f64(bk);
                }
            }
        }
        if (glane == 0 && live) {
            for (int c = 0; c < o.k; c++) {
                double r = 0.0;
                u8 ok = 0;
                bool mine = true;
                switch (o.ops[c]) {
                    case GRB_OP_SUM:
                        mine = sweep_sums;
                        r = sum;
                        ok = 1;
                        break;
                    case GRB_OP_SUM_NOT_EMPTY:
                        mine = sweep_sums;
                        ok = nvalid != 0;
                        r = ok ? sum : 0.0;
                        break;
                    case GRB_OP_MEAN:
                        mine = sweep_sums;
                        ok = nvalid != 0;
                        r = ok ? sum / (double)nvalid : 0.0;
                        break;
                    case GRB_OP_MIN:
                        ok = have;
                        r = have ? mn : 0.0;
                        break;
                    case GRB_OP_MAX:
                        ok = have;
                        r = have ? mx : 0.0;
                        break;
                    case GRB_OP_MEDIAN:
                        if (heavy) {
                            mine = false;
                        } else if (nvalid) {
                            ok = 1;
                            r = (nvalid & 1) ? med_lo : (med_lo + med_hi) / 2.0;
                        }
                        break;
                    case GRB_OP_OVERLAP_BP:
                        mine = WIDTHS;  // otherwise k_overlap_bp answers it in closed form
                        ok = 1;
                        r = (double)bp;
                        break;
                    case GRB_OP_WEIGHTED_MEAN:
                        mine = o.want_wm != 0;  // otherwise k_weighted_mean answers it in closed form
                        ok = 1;
                        r = wtot ? wsum / (double)wtot : 0.0;
                        break;
                    case GRB_OP_WEIGHTED_SUM:
                        mine = o.want_wm != 0;
                        ok = 1;
                        r = wsum;
                        break;
                    default:
                        mine = false;
                        break;
                }
                if (mine) {
                    o.out[g * (u64)o.k + c] = r;
                    o.out_valid[g * (u64)o.k + c] = ok;
                }
            }
            if (heavy) {
                const u32 slot = atomicAdd(o.heavy_count, 1u);
                o.heavy_list[slot] = (u32)g;  // the host checks total < 2^32 when a median is requested
            }
        }
        __syncwarp();  // the next left reuses kbuf
    }
}

// One CTA per tile of 1024 lefts (the tiling of k_tile).
__global__ void __launch_bounds__(SW2_TPB) k_sweep2(const SeqDev *__restrict__ seqs, const u32 *__restrict__ tile_begin, int nseq,
                                                    const u32 *__restrict__ ls, const u32 *__restrict__ le,
                                                    const u32 *__restrict__ ubv, const u32 *__restrict__ ppv,
                                                    const u32 *__restrict__ lbv, const SweepOut o) {
    __shared__ double s_buf[(SW2_TPB / 8) * MED2_CAP];  // 32 KB; viewed as [8][512] when G = 32
    __shared__ double s_med[(SW2_TPB / 8) * 2];
    __shared__ u32 s_cnt[SW2_TPB / 32];
    const u32 t = blockIdx.x;
    const SeqDev *__restrict__ sd = &seqs[find_seq_by_tile(tile_begin, nseq, t)];
    const u64 g_first = ((sd->left_begin >> TILE_SHIFT) + (t - sd->tile_begin)) << TILE_SHIFT;
    const u64 g_lo = sd->left_begin > g_first ? sd->left_begin : g_first;
    const u64 g_hi = sd->left_end < g_first + TILE ? sd->left_end : g_first + TILE;
    // members per left in this tile decide the group width (uniform for the CTA)
    u32 c = 0;
    for (u64 g = g_lo + threadIdx.x; g < g_hi; g += SW2_TPB) c += ubv[g] - lbv[g];
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    u64 total = 0;
#pragma unroll
    for (int w = 0; w < SW2_TPB / 32; w++) total += s_cnt[w];
    const bool wide_groups = total > 48ull * (g_hi - g_lo);
    if (o.want_bp) {
        if (wide_groups)
            sweep_tile<32, true>(sd, g_lo, g_hi, ls, le, ubv, ppv, lbv, o, s_buf, s_med);
        else
            sweep_tile<8, true>(sd, g_lo, g_hi, ls, le, ubv, ppv, lbv, o, s_buf, s_med);
    } else {
        if (wide_groups)
            sweep_tile<32, false>(sd, g_lo, g_hi, ls, le, ubv, ppv, lbv, o, s_buf, s_med);
        else
            sweep_tile<8, false>(sd, g_lo, g_hi, ls, le, ubv, ppv, lbv, o, s_buf, s_med);
    }
}

// Emit pass of left_overlaps: positions (start order) of the members of every group, ascending,
// which IS sort_ranges order (join.rs:121-128) because the index is sorted by (start, end, index).
// [pp, ub) is written as an iota without reading anything; the straddlers are found as above.
template <int G>
__device__ __forceinline__ void emit_tile(const SeqDev *__restrict__ sd, u64 g_lo, u64 g_hi, const u32 *__restrict__ ls,
                                          const u32 *__restrict__ ubv, const u32 *__restrict__ ppv, const u32 *__restrict__ lbv,
                                          const u64 *__restrict__ offsets, u32 *__restrict__ pos) {
    // Warp-synchronous like sweep3_run: the 32 / G groups of a warp step together and every vote is one
    // full-warp instruction (per-group masks make the hardware run each group's copy in turn).
    constexpr int NGROUPS = SW2_TPB / G;
    constexpr u32 GM = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const u32 glane = threadIdx.x & (G - 1), gidx = threadIdx.x / G;
    const u32 gshift = (threadIdx.x & 31) & ~(G - 1);
    const u32 *__restrict__ ends = sd->ends;
    const u32 *__restrict__ bme = sd->bme;
    const u64 g_warp = g_lo + (threadIdx.x >> 5) * (32 / G);
    const u32 rounds = g_warp < g_hi ? (u32)((g_hi - g_warp + NGROUPS - 1) / NGROUPS) : 0u;
    u64 g = g_lo + gidx;
    // Software pipeline, two lefts deep: while left i is written, the parameters of left i + 2 and the first straddler step of
    // left i + 1 (its block maximum and G ends, addressed by parameters that arrived a round ago) are in flight.  One left per
    // group and ~1 KB of pairs per left make the kernel a chain of dependent round trips otherwise: 43 % of its stall samples
    // sat on the load of the ends (config C5).
    auto first_step = [&](u32 p, u32 l, u32 &e0, u32 &bm0) {
        e0 = bm0 = 0;
        if (p > l && p) {
            const u32 cb = (p - 1) & ~(u32)(G - 1);
            const u32 j = cb + glane;
            if (j < p) e0 = ends[j];
            bm0 = bme[cb >> GRB_BME_SHIFT];
        }
    };
    u32 cu = 0, cp = 0, cl = 0, cls = 0, nu = 0, np = 0, nl = 0, nls = 0;
    u64 coff = 0, noff = 0;
    if (g < g_hi) {
        cu = ubv[g];
        cp = ppv[g];
        cl = lbv[g];
        cls = ls[g];
        coff = offsets[g];
    }
    if (g + NGROUPS < g_hi) {
        nu = ubv[g + NGROUPS];
        np = ppv[g + NGROUPS];
        nl = lbv[g + NGROUPS];
        nls = ls[g + NGROUPS];
        noff = offsets[g + NGROUPS];
    }
    u32 e0, bm0;
    first_step(cp, cl, e0, bm0);
    for (u32 round = 0; round < rounds; round++, g += NGROUPS) {
        const u32 u = cu, p = cp, l = cl, lstart = cls;
        u32 *__restrict__ dst = pos + coff;
        // in flight during this round: parameters two lefts ahead, first straddler step one left ahead
        u32 mu = 0, mp = 0, ml = 0, mls = 0;
        u64 moff = 0;
        if (g + 2 * NGROUPS < g_hi) {
            mu = ubv[g + 2 * NGROUPS];
            mp = ppv[g + 2 * NGROUPS];
            ml = lbv[g + 2 * NGROUPS];
            mls = ls[g + 2 * NGROUPS];
            moff = offsets[g + 2 * NGROUPS];
        }
        u32 ne0, nbm0;
        first_step(np, nl, ne0, nbm0);
        const u32 need = p - l;
        // [p, u) as an iota: a scalar head up to the next 16-byte boundary, 16-byte stores, a scalar tail
        {
            u32 *__restrict__ q = dst + need;
            const u32 cnt = u > p ? u - p : 0u;
            const u32 head = min(cnt, (4u - (u32)(((uintptr_t)q >> 2) & 3u)) & 3u);
            if (glane < head) q[glane] = p + glane;
            const u32 nvec = (cnt - head) >> 2;
            for (u32 v = glane; v < nvec; v += G) {
                const u32 j = p + head + 4u * v;
                *reinterpret_cast<uint4 *>(q + head + 4u * v) = make_uint4(j, j + 1, j + 2, j + 3);
            }
            const u32 t0 = head + 4u * nvec;
            if (t0 + glane < cnt) q[t0 + glane] = p + t0 + glane;
        }
        u32 found = 0;
        long long cb = p ? (long long)((p - 1) & ~(u32)(G - 1)) : -1;
        bool first = true;
        while (true) {
            const bool act = found < need && cb >= 0;
            if (!__any_sync(GRB_FULL, act)) break;
            const u32 j = (u32)cb + glane;
            u32 e, bm;
            if (first) {  // (warp-uniform: every group of the warp is in its first step)
                e = e0;
                bm = bm0;
                first = false;
            } else {
                e = act && j < p ? ends[j] : 0u;
                bm = act ? bme[cb >> GRB_BME_SHIFT] : 0u;
            }
            const bool look = act && bm > lstart;
            const bool hit = look && j < p && e > lstart;
            const u32 hm = (__ballot_sync(GRB_FULL, hit) >> gshift) & GM;
            if (hit) {
                const u32 above = glane == G - 1 ? 0u : __popc(hm >> (glane + 1));
                dst[need - 1 - (found + above)] = j;
            }
            found += __popc(hm);
            if (act) cb = look ? cb - G : ((cb >> GRB_BME_SHIFT) << GRB_BME_SHIFT) - G;
        }
        cu = nu, cp = np, cl = nl, cls = nls, coff = noff;
        nu = mu, np = mp, nl = ml, nls = mls, noff = moff;
        e0 = ne0, bm0 = nbm0;
    }
}

__global__ void __launch_bounds__(SW2_TPB) k_emit2(const SeqDev *__restrict__ seqs, const u32 *__restrict__ tile_begin, int nseq,
                                                   const u32 *__restrict__ ls, const u32 *__restrict__ ubv,
                                                   const u32 *__restrict__ ppv, const u32 *__restrict__ lbv,
                                                   const u64 *__restrict__ offsets, u32 *__restrict__ pos) {
    const u32 t = blockIdx.x;
    const SeqDev *__restrict__ sd = &seqs[find_seq_by_tile(tile_begin, nseq, t)];
    const u64 g_first = ((sd->left_begin >> TILE_SHIFT) + (t - sd->tile_begin)) << TILE_SHIFT;
    const u64 g_lo = sd->left_begin > g_first ? sd->left_begin : g_first;
    const u64 g_hi = sd->left_end < g_first + TILE ? sd->left_end : g_first + TILE;
    const u64 pairs = offsets[g_hi] - offsets[g_lo];
    // lanes per left.  The kernel is bound by its instructions per left (~190 per warp round whatever the group width), and a
    // lane writes 16 bytes per store: narrow groups share a round between more lefts (config C5, ~300 pairs per left:
    // 15.1 ms with a warp per left, EMIT_T16 / EMIT_T32 = average pairs per left from which 16 / 32 lanes take over)
#ifndef EMIT_T16
#define EMIT_T16 128ull
#endif
#ifndef EMIT_T32
#define EMIT_T32 1024ull
#endif
    if (pairs > EMIT_T32 * (g_hi - g_lo))
        emit_tile<32>(sd, g_lo, g_hi, ls, ubv, ppv, lbv, offsets, pos);
    else if (pairs > EMIT_T16 * (g_hi - g_lo))
        emit_tile<16>(sd, g_lo, g_hi, ls, ubv, ppv, lbv, offsets, pos);
    else
        emit_tile<8>(sd, g_lo, g_hi, ls, ubv, ppv, lbv, offsets, pos);
}
