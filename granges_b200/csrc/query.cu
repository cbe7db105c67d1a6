// query.cu -- the overlap-join hot path: count / filter / fused map reductions / pair emission.
//
// Reference path being replaced (per left range, src/granges.rs:958-969):
//     COITrees::query(start, end, |node| join.add_right(node))     ranges/coitrees.rs:48-57
//     JoinDataLeftEmpty::map -> gather right data by index          join.rs:343-361
//     drop None scores, FloatOperation::run per op                  commands.rs:524-542, operations.rs:53-109
//
// Here, per tile of 1024 consecutive lefts (one CTA):
//   1. four warp-cooperative 32-ary searches bound the slices of `starts` and `ends_sorted`
//      the tile can touch;
//   2. the two slices are staged into shared memory with TMA bulk copies (cp.async.bulk +
//      mbarrier) and every left resolves
//          ub = #{r.start < l.end}      lb = #{r.end <= l.start}
//      by a branch-free binary search in shared memory;  count = ub - lb  (exact);
//   3. SUM / MEAN come from exact fixed-point prefix sums PA[ub] - PB[lb] (index.cu), rounded
//      once to f64 -- no pair is visited;
//   4. MIN / MAX / MEDIAN / overlap bp need the members: a warp per left sweeps backward from
//      ub over `ends` (skipping 64-blocks whose block-max-end cannot overlap) until it has seen
//      `count` hits, reducing with ballots / shuffles; pair emission is the same sweep writing
//      positions behind a count -> scan -> emit CSR.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "tilelib.cuh"

namespace {

// exact 128-bit prefix of the scores at position j (start order: pfx_a / pfh_a / base_a, end order: pfx_b / pfh_b / base_b)
__device__ __forceinline__ i128 exact_prefix(const u64 *__restrict__ pfx, const u32 *__restrict__ pfh, const i128 *__restrict__ base, int mode,
                                             u32 j) {
    if (mode == GRB_SUM_WIDE) return base[j];
    const i128 b = base[j >> GRB_BASE_SHIFT];
    if (mode == GRB_SUM_WIDE96) {
        const u128 d = ((((u128)pfh[j] << 64) | pfx[j]) - (u128)b) & ((((u128)1) << 96) - 1);
        return b + sext96((u32)(d >> 64), (u64)d);
    }
    return b + (i128)(i64)(pfx[j] - (u64)(u128)b);
}

// ---- the tile kernel --------------------------------------------------------------
//
// One CTA = up to 1024 consecutive lefts of one seqname (tiles are aligned to multiples of 1024 in
// the concatenated left arrays, so full tiles read / write with 128-bit accesses); thread t owns
// lefts 4t..4t+3 of the tile.
//   a. load the lefts, block-reduce min/max of l.start and l.end;
//   b. warp 0 turns those into coordinate bins, reads the four bounding entries of the two
//      bin tables and issues TMA bulk copies of (starts slice, ends_sorted slice, both table
//      slices) into shared memory behind one mbarrier;
//   c. every left: bin -> [lo, hi] from the staged table -> short binary search in the staged keys;
//   d. outputs (counts / keep flags / ub,lb / fused reductions).
// FAST >= 0: the batch is uniform (every seqname has scores, no None, compact prefixes) and there
// is exactly one reduction, FAST = its grb_op code; the general path handles everything else.

template <int MODE, int FAST>
__global__ void __launch_bounds__(TILE_TPB) k_tile(const SeqDev *__restrict__ seqs, const u32 *__restrict__ tile_begin, int nseq,
                                                   const u32 *__restrict__ ls, const u32 *__restrict__ le, const u64 total,
                                                   const TileOut o, int use_tma) {
    __shared__ __align__(16) u32 sKA[WIN_CAP + 8];
    __shared__ __align__(16) u32 sKB[WIN_CAP + 8];
    __shared__ __align__(16) u32 sLA[LUT_CAP + 8];
    __shared__ __align__(16) u32 sLB[LUT_CAP + 8];
    __shared__ __align__(8) u64 mbar;
    __shared__ u32 s_red[4][TILE_TPB / 32];
    __shared__ u32 s_win[8];

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) mbar_init(&mbar, 1);

    const u32 t = blockIdx.x;
    const int s = find_seq_by_tile(tile_begin, nseq, t);
    const SeqDev *__restrict__ sd = &seqs[s];
    const u64 seq_lo = sd->left_begin, seq_hi = sd->left_end;
    const u64 gblock = (seq_lo >> TILE_SHIFT) + (t - sd->tile_begin);
    const u64 g0 = (gblock << TILE_SHIFT) + (u64)tid * TILE_LPT;  // first left of this thread (global index)
    const u32 n = sd->n;
    const u32 *__restrict__ starts = sd->starts;
    const u32 *__restrict__ ends_sorted = sd->ends_sorted;
    const u32 *__restrict__ lut_a = sd->lut_a;
    const u32 *__restrict__ lut_b = sd->lut_b;
    const int sh = sd->lut_shift;
    const u32 bmax = sd->lut_bmax;

    // ---- a. lefts
    u32 a[TILE_LPT], b[TILE_LPT];
    bool okv[TILE_LPT];
    const bool full = g0 >= seq_lo && g0 + TILE_LPT <= seq_hi;
    const bool vec_ok = ((((uintptr_t)ls) | ((uintptr_t)le)) & 15) == 0;
    if (full && vec_ok) {
        const uint4 va = *reinterpret_cast<const uint4 *>(ls + g0);
        const uint4 vb = *reinterpret_cast<const uint4 *>(le + g0);
        a[0] = va.x, a[1] = va.y, a[2] = va.z, a[3] = va.w;
        b[0] = vb.x, b[1] = vb.y, b[2] = vb.z, b[3] = vb.w;
#pragma unroll
        for (int k = 0; k < TILE_LPT; k++) okv[k] = true;
    } else {
#pragma unroll
        for (int k = 0; k < TILE_LPT; k++) {
            const u64 g = g0 + k;
            okv[k] = g >= seq_lo && g < seq_hi;
            a[k] = okv[k] ? ls[g] : 0u;
            b[k] = okv[k] ? le[g] : 0u;
        }
    }
    u32 mn_s = 0xffffffffu, mx_s = 0, mn_e = 0xffffffffu, mx_e = 0;
    // a left with end <= start or end > 2^31-1 cannot exist in the reference (it panics); flag it and answer it
    // as (0, 0), which overlaps nothing, so every later index stays in range
    bool bad = false;
#pragma unroll
    for (int k = 0; k < TILE_LPT; k++) {
        if (okv[k] && (b[k] <= a[k] || b[k] > 0x7fffffffu)) {
            bad = true;
            a[k] = b[k] = 0;
        }
    }
    if (__any_sync(GRB_FULL, bad) && lane == 0) atomicOr(o.flags, 1u);
#pragma unroll
    for (int k = 0; k < TILE_LPT; k++) {
        if (okv[k]) {
            mn_s = min(mn_s, a[k]);
            mx_s = max(mx_s, a[k]);
            mn_e = min(mn_e, b[k]);
            mx_e = max(mx_e, b[k]);
        }
    }
    mn_e = warp_min(mn_e);
    mx_e = warp_max(mx_e);
    mn_s = warp_min(mn_s);
    mx_s = warp_max(mx_s);
    if (lane == 0) {
        s_red[0][warp] = mn_e;
        s_red[1][warp] = mx_e;
        s_red[2][warp] = mn_s;
        s_red[3][warp] = mx_s;
    }
    __syncthreads();

    // ---- b. windows (warp 0)
    if (warp == 0) {
        // lane q in 0..3 finishes reduction q (min for even q, max for odd q) and fetches one table entry
        const u32 q = lane & 3;
        // with M_UBLB the starts window must also cover the searches for l.start: its low end comes from min l.start
        const u32 row = ((MODE & M_UBLB) && q == 0) ? 2u : q;
        u32 v = (q & 1) ? 0u : 0xffffffffu;
#pragma unroll
        for (int w = 0; w < TILE_TPB / 32; w++) {
            u32 p = s_red[row][w];
            v = (q & 1) ? max(v, p) : min(v, p);
        }
        // x = l.end for the starts table (q = 0,1); x = l.start + 1 for the ends table (q = 2,3)
        const u32 x = v + (q >> 1);
        const u32 bin = min(x >> sh, bmax);
        const u32 *tab = (q < 2) ? lut_a : lut_b;
        const u32 bound = tab[bin + (q & 1)];  // lower bins give the window start, upper bin + 1 its end
        const u32 binA0 = __shfl_sync(GRB_FULL, bin, 0), binA1 = __shfl_sync(GRB_FULL, bin, 1);
        const u32 binB0 = __shfl_sync(GRB_FULL, bin, 2), binB1 = __shfl_sync(GRB_FULL, bin, 3);
        const u32 A_lo = __shfl_sync(GRB_FULL, bound, 0), A_hi = __shfl_sync(GRB_FULL, bound, 1);
        const u32 B_lo = __shfl_sync(GRB_FULL, bound, 2), B_hi = __shfl_sync(GRB_FULL, bound, 3);
        if (lane == 0) {
            const u32 a0 = A_lo & ~3u, b0 = B_lo & ~3u, la0 = binA0 & ~3u, lb0 = binB0 & ~3u;
            // + 4: the forward scan may read up to 3 keys past the last one it can return
            const u32 alen = A_hi - a0 + 4, blen = B_hi - b0 + 4;
            const u32 lalen = binA1 + 1 - la0, lblen = binB1 + 1 - lb0;
            const bool staged = use_tma && alen <= WIN_CAP && blen <= WIN_CAP && lalen <= LUT_CAP && lblen <= LUT_CAP;
            s_win[0] = a0;
            s_win[1] = b0;
            s_win[2] = la0;
            s_win[3] = lb0;
            s_win[4] = staged ? 1u : 0u;
            if (staged) {
                const u32 by_a = (alen * 4 + 15) & ~15u, by_b = (blen * 4 + 15) & ~15u;
                const u32 by_la = (lalen * 4 + 15) & ~15u, by_lb = (lblen * 4 + 15) & ~15u;
                mbar_arrive_expect_tx(&mbar, by_a + by_b + by_la + by_lb);
                tma_load_1d(sLA, lut_a + la0, by_la, &mbar);
                tma_load_1d(sLB, lut_b + lb0, by_lb, &mbar);
                tma_load_1d(sKA, starts + a0, by_a, &mbar);
                tma_load_1d(sKB, ends_sorted + b0, by_b, &mbar);
            }
        }
    }
    __syncthreads();
    const u32 a0 = s_win[0], b0 = s_win[1], la0 = s_win[2], lb0 = s_win[3];
    const bool staged = s_win[4] != 0;

    // ---- c. per-left searches
    u32 ub[TILE_LPT], lb[TILE_LPT], pp[TILE_LPT];
    if (staged) {
        mbar_wait(&mbar, 0);
#pragma unroll
        for (int k = 0; k < TILE_LPT; k++) {
            ub[k] = lb[k] = pp[k] = 0;
            if (!okv[k]) continue;
            const u32 xa = b[k], xb = a[k] + 1;
            const u32 ba = min(xa >> sh, bmax) - la0;
            const u32 bb = min(xb >> sh, bmax) - lb0;
            ub[k] = a0 + scan_lower_bound(sKA, sLA[ba] - a0, xa);
            lb[k] = b0 + scan_lower_bound(sKB, sLB[bb] - b0, xb);
            if (MODE & M_UBLB) pp[k] = a0 + scan_lower_bound(sKA, sLA[min(a[k] >> sh, bmax) - la0] - a0, a[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < TILE_LPT; k++) {
            const u32 xa = b[k], xb = a[k] + 1;
            const u32 ba = min(xa >> sh, bmax), bb = min(xb >> sh, bmax);
            ub[k] = okv[k] ? scan_lower_bound(starts, lut_a[ba], xa) : 0u;
            lb[k] = okv[k] ? scan_lower_bound(ends_sorted, lut_b[bb], xb) : 0u;
            pp[k] = 0;
            if ((MODE & M_UBLB) && okv[k]) pp[k] = scan_lower_bound(starts, lut_a[min(a[k] >> sh, bmax)], a[k]);
        }
    }

    // ---- d. outputs
    if (FAST >= 0) {
        // uniform batch, one reduction, k == 1
        const u64 *__restrict__ pfx_a = sd->pfx_a;
        const u64 *__restrict__ pfx_b = sd->pfx_b;
        const u32 fast_cmax = sd->fast_cmax;
        const double scale = sd->scale;
        double r[TILE_LPT];
        u8 v[TILE_LPT];
#pragma unroll
        for (int k = 0; k < TILE_LPT; k++) {
            const u32 c = ub[k] - lb[k];
            if (FAST == GRB_OP_COUNT) {
                r[k] = (double)c;
                v[k] = 1;
            } else {
                double sum = 0.0;
                if (okv[k]) sum = compact_group_sum(pfx_a, pfx_b, sd->base_a, sd->base_b, ub[k], lb[k], fast_cmax, scale, sd->e0);
                if (FAST == GRB_OP_SUM) {
                    r[k] = sum;
                    v[k] = 1;
                } else if (FAST == GRB_OP_SUM_NOT_EMPTY) {
                    v[k] = c != 0;
                    r[k] = sum;
                } else {
                    v[k] = c != 0;
                    r[k] = c ? sum / (double)c : 0.0;
                }
            }
        }
        const bool ovec = ((((uintptr_t)o.out) & 15) | (((uintptr_t)o.out_valid) & 3)) == 0;
        if (full && ovec) {
            double2 *po = reinterpret_cast<double2 *>(o.out + g0);
            po[0] = make_double2(r[0], r[1]);
            po[1] = make_double2(r[2], r[3]);
            *reinterpret_cast<uchar4 *>(o.out_valid + g0) = make_uchar4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < TILE_LPT; k++) {
                if (okv[k]) {
                    o.out[g0 + k] = r[k];
                    o.out_valid[g0 + k] = v[k];
                }
            }
        }
        if (MODE & M_UBLB) {
#pragma unroll
            for (int k = 0; k < TILE_LPT; k++) {
                if (okv[k]) {
                    o.ub[g0 + k] = ub[k];
                    o.lb[g0 + k] = lb[k];
                    o.pp[g0 + k] = pp[k];
                }
            }
        }
        return;
    }

    const int mode = sd->sum_mode;
    const int has_nulls = sd->has_nulls, has_scores = sd->has_scores;
    u32 kept = 0;
#pragma unroll
    for (int k = 0; k < TILE_LPT; k++) {
        if (!okv[k]) continue;
        const u64 g = g0 + k;
        const u32 c_all = ub[k] - lb[k];
        if (MODE & M_COUNTS) o.counts[g] = c_all;
        if (MODE & M_KEEP) {
            u8 kp = (u8)((c_all != 0) != (o.anti != 0));
            o.keep[g] = kp;
            kept += kp;
        }
        if (MODE & M_UBLB) {
            o.ub[g] = ub[k];
            o.lb[g] = lb[k];
            o.pp[g] = pp[k];
        }
        if (MODE & M_MAP) {
            u32 nvalid = c_all;
            if (has_nulls) nvalid = sd->vcnt_a[ub[k]] - sd->vcnt_b[lb[k]];
            if (!has_scores) nvalid = 0;
            double sum = 0.0;
            if (o.need_sum && mode != GRB_SUM_SWEEP && has_scores) {
                sum = fx_to_double(exact_prefix(sd->pfx_a, sd->pfh_a, sd->base_a, mode, ub[k]) -
                                       exact_prefix(sd->pfx_b, sd->pfh_b, sd->base_b, mode, lb[k]),
                                   sd->e0);
                if (sd->nf_a) {
                    const uint4 na = sd->nf_a[ub[k]], nb = sd->nf_b[lb[k]];
                    sum = nf_rule(sum, na.x - nb.x, na.y - nb.y, na.z - nb.z);
                }
            }
            for (int c = 0; c < o.k; c++) {
                const int op = o.ops[c];
                double r = 0.0;
                u8 ok = 0;
                bool mine = true;
                switch (op) {
                    case GRB_OP_COUNT:
                        r = (double)c_all;
                        ok = 1;
                        break;
                    case GRB_OP_SUM:
                        mine = mode != GRB_SUM_SWEEP;
                        r = sum;
                        ok = 1;
                        break;
                    case GRB_OP_SUM_NOT_EMPTY:
                        mine = mode != GRB_SUM_SWEEP;
                        ok = nvalid != 0;
                        r = ok ? sum : 0.0;
                        break;
                    case GRB_OP_MEAN:
                        mine = mode != GRB_SUM_SWEEP;
                        ok = nvalid != 0;
                        r = ok ? sum / (double)nvalid : 0.0;
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
        }
    }
    if (MODE & M_KEEP) {
        if (o.n_kept) {
            kept = warp_sum(kept);
            if (lane == 0 && kept) atomicAdd(o.n_kept, (unsigned long long)kept);
        }
    }
}

// ---- the sweep kernels ----------------------------------------------------------------

constexpr int SW_TPB = 256;

struct SweepOut {
    double *out;
    u8 *out_valid;
    int k;
    int ops[GRB_MAX_OPS];
    int want_median, want_bp, want_wm;  // want_wm implies want_bp (the overlap widths are needed)
    u32 *heavy_list;
    u32 *heavy_count;
    const uint4 *planar;  // k_sweep3: (sum, count, scored members) per left from the pipelined pass: the prefix-sum columns are written here, with the rows
    int wave_ratio;  // k_sweep3: a run of lefts answers its medians from per-run wavelet trees when lefts * wave_ratio >= rights of its slice (0: never)
};

__device__ __forceinline__ bool score_valid(const SeqDev *sd, u32 j) {
    if (!sd->has_scores) return false;
    if (!sd->has_nulls) return true;
    return (sd->vbits_a[j >> 5] >> (j & 31)) & 1u;
}

// Median of groups too large for the in-group quickselect of k_sweep2: one CTA per such left, MSB-first radix
// select (8 bits per pass) over the order-preserving integer image of the scores, re-sweeping the
// candidates instead of materialising them.
__global__ void __launch_bounds__(SW_TPB) k_median_heavy(const SeqDev *__restrict__ seqs, int nseq, const u32 *__restrict__ ls,
                                                         const u32 *__restrict__ ubv, const u32 *__restrict__ lbv,
                                                         const u32 *__restrict__ heavy_list, const u32 *__restrict__ heavy_count,
                                                         const SweepOut o) {
    __shared__ u32 hist[256];
    __shared__ u64 s_prefix;
    __shared__ u32 s_kk;
    const u32 nheavy = *heavy_count;
    for (u32 h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const u64 g = heavy_list[h];
        const SeqDev *sd = &seqs[find_seq_by_left(seqs, nseq, g)];
        const u32 u = ubv[g], count = u - lbv[g], lstart = ls[g];
        const u32 *__restrict__ ends = sd->ends;
        // 1. how far back do the members reach, and how many carry a score?
        u32 found = 0, m = 0, j_top = u;
        while (found < count && j_top > 0) {
            long long j = (long long)j_top - 1 - threadIdx.x;
            bool hit = j >= 0 && ends[j] > lstart;
            bool v = hit && score_valid(sd, (u32)j);
            found += __syncthreads_count(hit);
            m += __syncthreads_count(v);
            j_top = j_top > SW_TPB ? j_top - SW_TPB : 0;
        }
        const u32 j_lo = j_top;
        double res[2] = {0.0, 0.0};
        const u32 targets[2] = {(m - 1) >> 1, m >> 1};
        for (int which = 0; which < 2; which++) {
            if (which == 1 && targets[1] == targets[0]) {
                res[1] = res[0];
                break;
            }
            if (threadIdx.x == 0) {
                s_prefix = 0;
                s_kk = targets[which];
            }
            for (int pass = 7; pass >= 0; pass--) {
                hist[threadIdx.x] = 0;
                __syncthreads();
                const u64 prefix = s_prefix;
                for (u32 j = j_lo + threadIdx.x; j < u; j += SW_TPB) {
                    if (ends[j] > lstart && score_valid(sd, j)) {
                        u64 key = f64_key(sd->score_a[j]);
                        bool match = pass == 7 ? true : ((key ^ prefix) >> (8 * (pass + 1))) == 0;
                        if (match) atomicAdd(&hist[(key >> (8 * pass)) & 0xff], 1u);
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    u32 kk = s_kk, cum = 0;
                    int d = 0;
                    for (; d < 255; d++) {
                        if (cum + hist[d] > kk) break;
                        cum += hist[d];
                    }
                    s_kk = kk - cum;
                    s_prefix = prefix | ((u64)d << (8 * pass));
                }
                __syncthreads();
            }
            res[which] = key_f64(s_prefix);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            for (int c = 0; c < o.k; c++) {
                if (o.ops[c] != GRB_OP_MEDIAN) continue;
                o.out[g * (u64)o.k + c] = (m & 1) ? res[1] : (res[0] + res[1]) / 2.0;
                o.out_valid[g * (u64)o.k + c] = 1;
            }
        }
        __syncthreads();
    }
}

__global__ void k_pairs_gather(const u32 *__restrict__ pos, u64 np, const u32 *__restrict__ starts, const u32 *__restrict__ ends,
                               const u32 *__restrict__ didx, u64 *__restrict__ out_idx, u32 *__restrict__ out_s,
                               u32 *__restrict__ out_e) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    u32 j = pos[i];
    if (out_idx) out_idx[i] = didx[j];
    if (out_s) out_s[i] = starts[j];
    if (out_e) out_e[i] = ends[j];
}

// overlap_width (traits.rs:73-80) of every pair: one warp per left walks its group.
__global__ void k_pairs_widths(const u32 *__restrict__ pos, const u64 *__restrict__ offsets, u64 nl, const u32 *__restrict__ ls,
                               const u32 *__restrict__ le, const u32 *__restrict__ starts, const u32 *__restrict__ ends,
                               u32 *__restrict__ widths) {
    const u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= nl) return;
    const u32 a = ls[g], b = le[g];
    for (u64 i = offsets[g] + lane_id(); i < offsets[g + 1]; i += 32) {
        u32 j = pos[i];
        u32 s = max(a, starts[j]), e = min(b, ends[j]);
        widths[i] = e > s ? e - s : 0u;
    }
}

#include "sweep.cuh"
#include "sweep3.cuh"

// OVERLAP_BP (sum of LeftGroupedJoin::overlap_widths, join.rs:158-163) in closed form from ub / pp / lb and the
// prefix sums of the sorted starts / ends (index.cu: grb_index_prepare_bp): one more bounded search per left.
__global__ void __launch_bounds__(256) k_overlap_bp(const SeqDev *__restrict__ seqs, int nseq, u64 first, u64 total, const u32 *__restrict__ ls,
                                                    const u32 *__restrict__ le, const u32 *__restrict__ ubv,
                                                    const u32 *__restrict__ ppv, const u32 *__restrict__ lbv,
                                                    double *__restrict__ out, u8 *__restrict__ out_valid, int k, u32 col_mask) {
    const u64 g = first + (u64)blockIdx.x * blockDim.x + threadIdx.x;  // a batch may start anywhere in the left arrays
    if (g >= total) return;
    const SeqDev *__restrict__ sd = &seqs[find_seq_by_left(seqs, nseq, g)];
    u64 bp = 0;
    if (sd->n) {
        const u32 u = ubv[g], p = ppv[g], l = lbv[g], a = ls[g], b = le[g];
        // #{end <= l.end} lies in [lb, ub]: an end <= l.start is <= l.end, and end <= l.end implies start < l.end
        u32 lo = l, hi = u;
        if (hi < lo) hi = lo;  // invalid left (flagged elsewhere)
        const u32 *__restrict__ es = sd->ends_sorted;
        while (lo < hi) {
            const u32 mid = lo + ((hi - lo) >> 1);
            if (es[mid] <= b)
                lo = mid + 1;
            else
                hi = mid;
        }
        const u32 a2 = lo;
        const u64 f_end = sd->sum_ends[a2] - sd->sum_starts[u] + (u64)b * (u64)(u - a2);
        const u64 f_start = sd->sum_ends[l] - sd->sum_starts[p] + (u64)a * (u64)(p - l);
        bp = f_end - f_start;
    }
    for (int c = 0; c < k; c++)
        if ((col_mask >> c) & 1u) {
            out[g * (u64)k + c] = (double)bp;
            out_valid[g * (u64)k + c] = 1;
        }
}

// MIN / MAX without a sweep (index.cu: grb_index_prepare_minmax): the straddlers of l.start from the stabbing tables at
// t = pp + lb, the members that start inside the left from the block sparse table over [pp, ub) -- at most 62 ranks are
// read one by one (the partial blocks at both ends), everything in between is two table cells.
__global__ void __launch_bounds__(256) k_minmax(const SeqDev *__restrict__ seqs, int nseq, u64 first, u64 total,
                                                const u32 *__restrict__ ubv, const u32 *__restrict__ ppv, const u32 *__restrict__ lbv,
                                                double *__restrict__ out, u8 *__restrict__ out_valid, int k, u32 min_cols, u32 max_cols) {
    const u64 g = first + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const SeqDev *__restrict__ sd = &seqs[find_seq_by_left(seqs, nseq, g)];
    u32 mn = GRB_RANK_NONE;
    int mx = -1;
    if (sd->n && sd->has_scores) {
        const u32 u = ubv[g], p = ppv[g], l = lbv[g];
        const u32 t = p + l;
        mn = sd->stab_min[t];
        mx = sd->stab_max[t];
        if (u > p) {
            const u32 *__restrict__ rk = sd->rank_a;
            const u32 fa = p >> 5, la = (u - 1) >> 5;
            if (la - fa <= 1) {
                for (u32 j = p; j < u; j++) {
                    const u32 r = rk[j];
                    mn = min(mn, r);
                    mx = max(mx, (int)r);
                }
            } else {
                for (u32 j = p; j < ((fa + 1) << 5); j++) {
                    const u32 r = rk[j];
                    mn = min(mn, r);
                    mx = max(mx, (int)r);
                }
                for (u32 j = la << 5; j < u; j++) {
                    const u32 r = rk[j];
                    mn = min(mn, r);
                    mx = max(mx, (int)r);
                }
                const u32 b0 = fa + 1, b1 = la;  // whole blocks [b0, b1)
                const int kk = 31 - __clz(b1 - b0);
                const size_t base = (size_t)kk * sd->sp_nb;
                mn = min(mn, min(sd->sp_min[base + b0], sd->sp_min[base + b1 - (1u << kk)]));
                mx = max(mx, max(sd->sp_max[base + b0], sd->sp_max[base + b1 - (1u << kk)]));
            }
        }
    }
    const double vmin = mn != GRB_RANK_NONE ? sd->score_sorted[mn] : 0.0;
    const double vmax = mx >= 0 ? sd->score_sorted[mx] : 0.0;
    for (int c = 0; c < k; c++) {
        if ((min_cols >> c) & 1u) {
            out[g * (u64)k + c] = vmin;
            out_valid[g * (u64)k + c] = mn != GRB_RANK_NONE;
        } else if ((max_cols >> c) & 1u) {
            out[g * (u64)k + c] = vmax;
            out_valid[g * (u64)k + c] = mx >= 0;
        }
    }
}

// WEIGHTED_MEAN (examples/weighted_mean.rs:4-31: sum(score * overlap width) / sum(overlap width) over the rights that
// carry a score, 0.0 if there is none) in closed form, like k_overlap_bp: no member is visited and the weighted sum is
// exact before it is rounded once.
__global__ void __launch_bounds__(256) k_weighted_mean(const SeqDev *__restrict__ seqs, int nseq, u64 first, u64 total, const u32 *__restrict__ ls,
                                                       const u32 *__restrict__ le, const u32 *__restrict__ ubv,
                                                       const u32 *__restrict__ ppv, const u32 *__restrict__ lbv,
                                                       double *__restrict__ out, u8 *__restrict__ out_valid, int k, u32 col_mask,
                                                       u32 sum_mask /* columns that want the weighted SUM (the numerator) */) {
    const u64 g = first + (u64)blockIdx.x * blockDim.x + threadIdx.x;  // a batch may start anywhere in the left arrays
    if (g >= total) return;
    const SeqDev *__restrict__ sd = &seqs[find_seq_by_left(seqs, nseq, g)];
    double r = 0.0, rsum = 0.0;
    if (sd->n && sd->has_scores) {
        const u32 u = ubv[g], p = ppv[g], l = lbv[g], a = ls[g], b = le[g];
        u32 lo = l, hi = u;
        if (hi < lo) hi = lo;  // invalid left (flagged elsewhere)
        const u32 *__restrict__ es = sd->ends_sorted;
        while (lo < hi) {
            const u32 mid = lo + ((hi - lo) >> 1);
            if (es[mid] <= b)
                lo = mid + 1;
            else
                hi = mid;
        }
        const u32 a2 = lo;  // #{end <= l.end}
        const int mode = sd->sum_mode;
        const i128 pa_u = exact_prefix(sd->pfx_a, sd->pfh_a, sd->base_a, mode, u), pa_p = exact_prefix(sd->pfx_a, sd->pfh_a, sd->base_a, mode, p);
        const i128 pb_a2 = exact_prefix(sd->pfx_b, sd->pfh_b, sd->base_b, mode, a2), pb_l = exact_prefix(sd->pfx_b, sd->pfh_b, sd->base_b, mode, l);
        const i128 g_end = sd->wsum_b[a2] - sd->wsum_a[u] + (i128)b * (pa_u - pb_a2);
        const i128 g_start = sd->wsum_b[l] - sd->wsum_a[p] + (i128)a * (pa_p - pb_l);
        const i128 wsum = g_end - g_start;
        u64 wtot;
        if (sd->has_nulls) {
            const u32 vu = sd->vcnt_a[u], vp = sd->vcnt_a[p], va2 = sd->vcnt_b[a2], vl = sd->vcnt_b[l];
            wtot = (sd->sum_ends_v[a2] - sd->sum_starts_v[u] + (u64)b * (u64)(vu - va2)) -
                   (sd->sum_ends_v[l] - sd->sum_starts_v[p] + (u64)a * (u64)(vp - vl));
        } else {
            wtot = (sd->sum_ends[a2] - sd->sum_starts[u] + (u64)b * (u64)(u - a2)) -
                   (sd->sum_ends[l] - sd->sum_starts[p] + (u64)a * (u64)(p - l));
        }
        rsum = fx_to_double(wsum, sd->e0);
        if (wtot) r = rsum / (double)wtot;
    }
    for (int c = 0; c < k; c++)
        if (((col_mask | sum_mask) >> c) & 1u) {
            out[g * (u64)k + c] = ((sum_mask >> c) & 1u) ? rsum : r;
            out_valid[g * (u64)k + c] = 1;
        }
}

// filter_overlaps as a stream compaction: keep flags -> exclusive scan -> scatter of the kept left indices
// (ascending, i.e. the reference's order-preserving filter, src/granges.rs:1485-1508).
__global__ void k_keep_to_u32(const u8 *__restrict__ keep, u64 n, u32 *__restrict__ flags) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = keep[i] ? 1u : 0u;
}
__global__ void k_scatter_kept(const u8 *__restrict__ keep, const u64 *__restrict__ pos, u64 n, u64 *__restrict__ kept_index) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keep[i]) kept_index[pos[i]] = i;
}

// ---- host orchestration -------------------------------------------------------------------

int make_batch(grb_ctx *ctx, const grb_index *const *idx, const u64 *left_offsets, int nseq, Batch *b) {
    if (nseq <= 0) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "batch: nseq must be >= 1");
    std::vector<SeqDev> h((size_t)nseq);
    std::vector<u32> tb((size_t)nseq + 1);
    u64 tiles = 0;
    b->pipeable = true;
    b->uniform = true;
    b->var.assign((size_t)nseq, -1);
    const u32 *zeros = (const u32 *)ctx->zeros;
    for (int s = 0; s < nseq; s++) {
        if (left_offsets[s + 1] < left_offsets[s]) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "batch: left_offsets must be non-decreasing");
        SeqDev d;
        memset(&d, 0, sizeof(d));
        const grb_index *ix = idx ? idx[s] : nullptr;
        if (ix && ix->ctx != ctx) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "batch: index belongs to another context");
        if (ix && ix->n) {
            d.starts = ix->starts;
            d.ends = ix->ends;
            d.ends_sorted = ix->ends_sorted;
            d.didx = ix->didx;
            d.bme = ix->bme;
            d.score_a = ix->score_a;
            d.vbits_a = ix->vbits_a;
            d.vcnt_a = ix->vcnt_a;
            d.vcnt_b = ix->vcnt_b;
            d.pfx_a = ix->pfx_a;
            d.pfx_b = ix->pfx_b;
            d.pfh_a = ix->pfh_a;
            d.pfh_b = ix->pfh_b;
            d.nf_a = ix->nf_a;
            d.nf_b = ix->nf_b;
            d.vc16_a = ix->vc16_a;
            d.vc16_b = ix->vc16_b;
            d.lut_a = ix->lut_a;
            d.lut_b = ix->lut_b;
            d.lut16_a = ix->lut16_a;
            d.lut16_b = ix->lut16_b;
            d.rank_a = ix->rank_a;
            d.perm_e = ix->perm_e;
            d.score_sorted = ix->score_sorted;
            d.fb = ix->fb;
            d.stab_min = ix->stab_min;
            d.stab_max = ix->stab_max;
            d.sp_min = ix->sp_min;
            d.sp_max = ix->sp_max;
            d.sp_nb = ix->sp_nb;
            d.sum_starts = ix->sum_starts;
            d.sum_ends = ix->sum_ends;
            d.sum_starts_v = ix->sum_starts_v;
            d.sum_ends_v = ix->sum_ends_v;
            d.wsum_a = ix->wsum_a;
            d.wsum_b = ix->wsum_b;
            d.lut_shift = ix->lut_shift;
            d.lut_bmax = ix->lut_bmax;
            d.base_a = ix->base_a;
            d.base_b = ix->base_b;
            d.n = ix->n;
            d.e0 = ix->e0;
            d.scale = ldexp(1.0, ix->e0);
            d.sum_mode = ix->has_scores ? ix->sum_mode : 3;
            d.has_nulls = ix->has_nulls;
            d.has_scores = ix->has_scores;
            d.fast_cmax = 0;
            if (d.sum_mode == GRB_SUM_COMPACT || d.sum_mode == GRB_SUM_WIDE96) {
                // a group of fewer than 2^room members sums to < 2^63 (2^95): the wrapped prefixes suffice
                int room = (d.sum_mode == GRB_SUM_COMPACT ? 63 : 95) - ix->width;
                d.fast_cmax = room >= 31 ? 0x7fffffffu : (1u << room);
            }
            if (!(ix->has_scores && !ix->has_nulls && !ix->has_nonfinite && ix->sum_mode == GRB_SUM_COMPACT)) b->uniform = false;
            if (ix->has_scores && (ix->sum_mode == GRB_SUM_COMPACT || ix->sum_mode == GRB_SUM_WIDE96))
                b->var[s] = ((ix->has_nulls || ix->has_nonfinite) ? PIPE_VAR_NULLS : 0) | (ix->sum_mode == GRB_SUM_WIDE96 ? PIPE_VAR_WIDE : 0);
            else
                b->pipeable = false;  // no scores / full 128-bit prefixes / member sums: the tile kernel (and the sweep)
        } else {
            // absent seqname / empty index: every group is empty; all tables read as zeros
            d.lut_a = d.lut_b = zeros;
            d.lut16_a = d.lut16_b = d.vc16_a = d.vc16_b = (const u16 *)zeros;
            d.pfh_a = d.pfh_b = zeros;
            d.starts = d.ends = d.ends_sorted = zeros + 256;  // sentinel keys: every forward scan stops at once
            d.pfx_a = d.pfx_b = (const u64 *)zeros;
            d.base_a = d.base_b = (const i128 *)zeros;
            d.lut_shift = 31;
            d.lut_bmax = 0;
            d.scale = 1.0;
            d.fast_cmax = 0x7fffffffu;
            d.sum_mode = 3;  // nothing to sum
        }
        d.left_begin = left_offsets[s];
        d.left_end = left_offsets[s + 1];
        d.tile_begin = (u32)tiles;
        tb[s] = (u32)tiles;
        // tiles are aligned to multiples of TILE in the concatenated left index space
        if (d.left_end > d.left_begin) tiles += ((d.left_end + TILE - 1) >> TILE_SHIFT) - (d.left_begin >> TILE_SHIFT);
        if (tiles >= 0x7fffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "batch: too many left ranges for one call");
        h[s] = d;
    }
    tb[nseq] = (u32)tiles;
    b->host_tile_begin = tb;
    b->ntiles = (u32)tiles;
    b->total = left_offsets[nseq];
    b->nseq = nseq;
    // descriptors: [SeqDev x nseq][u32 x (nseq + 1)] in one device buffer owned by the context; a call with the
    // same batch as the previous one (the usual case when a query is repeated) reuses it without any copy
    const size_t seq_bytes = sizeof(SeqDev) * nseq, bytes = seq_bytes + sizeof(u32) * (nseq + 1);
    std::vector<unsigned char> blob(bytes);
    memcpy(blob.data(), h.data(), seq_bytes);
    memcpy(blob.data() + seq_bytes, tb.data(), sizeof(u32) * (nseq + 1));
    if (!(ctx->batch_dev && ctx->batch_bytes == bytes && memcmp(ctx->batch_host, blob.data(), bytes) == 0)) {
        if (bytes > ctx->batch_cap) {
            GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->batch_dev);
            free(ctx->batch_host);
            ctx->batch_dev = nullptr;
            ctx->batch_cap = 0;
            ctx->batch_bytes = 0;
            ctx->batch_host = malloc(bytes * 2);
            if (!ctx->batch_host) return grb_fail(ctx, GRB_ERR_OOM, "batch: host allocation failed");
            GRB_CUDA(ctx, cudaMalloc(&ctx->batch_dev, bytes * 2));
            ctx->batch_cap = bytes * 2;
        }
        // The copy below is ordered on the stream after every kernel that still reads the previous batch, and
        // batch_host is pageable, so cudaMemcpyAsync has staged it by the time it returns: no synchronisation.
        memcpy(ctx->batch_host, blob.data(), bytes);
        ctx->batch_bytes = bytes;
        GRB_CUDA(ctx, cudaMemcpyAsync(ctx->batch_dev, ctx->batch_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    b->seqs = (const SeqDev *)ctx->batch_dev;
    b->tile_begin = (const u32 *)((const unsigned char *)ctx->batch_dev + seq_bytes);
    return GRB_OK;
}

// Which pipelined kernel answers a plain map call (no member reduction, no closed form): the op code of a single
// reduction, PIPE_MULTI for several, -1 when the batch or the reductions need the tile kernel.
int map_pipe_fast(const Batch &b, const TileOut &o) {
    bool need_scores = false;
    for (int c = 0; c < o.k; c++) {
        const int op = o.ops[c];
        if (op != GRB_OP_SUM && op != GRB_OP_SUM_NOT_EMPTY && op != GRB_OP_MEAN && op != GRB_OP_COUNT) return -1;
        need_scores |= op != GRB_OP_COUNT;
    }
    if (need_scores && !b.pipeable) return -1;
    return o.k == 1 ? o.ops[0] : PIPE_MULTI;
}

template <int MODE, int FAST>
int launch_tile(grb_ctx *ctx, const Batch &b, const u32 *ls, const u32 *le, const TileOut &o) {
    if (b.ntiles == 0) return GRB_OK;
    k_tile<MODE, FAST><<<b.ntiles, TILE_TPB, 0, ctx->stream>>>(b.seqs, b.tile_begin, b.nseq, ls, le, b.total, o, ctx->tile_mode);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

// MAP launch: the pipelined kernel when the prefix sums of every seqname can be staged, else the tile kernels.
// valid_bits (single-reduction pipelined launches only; the caller checked): validity as bits instead of bytes.
template <int MODE>
int launch_map(grb_ctx *ctx, const Batch &b, const u32 *ls, const u32 *le, const TileOut &o, u32 *valid_bits = nullptr) {
    const bool ublb_pipe = ctx->ublb_pipe < 0 ? o.need_sum != 0 : ctx->ublb_pipe != 0;
    if (MODE == (M_MAP | M_UBLB) && ctx->pipe_mode == 1 && ublb_pipe && (b.pipeable || !o.need_sum) && b.ntiles >= 4u * (u32)ctx->sm_count) {
        // The counts for the member sweep / the closed forms, with the columns the prefix sums answer, in the pipeline.
        // Measured at 100M x 100M: with sums beside the members (config C3) the pipeline saves 0.8 ms; for counts alone
        // the tile kernel is as fast (3.1 vs 3.3 ms for `min`), so it keeps those.
        return grb_launch_pipe(ctx, b, ls, le, o, PIPE_UBLB, nullptr);
    }
    if (MODE == M_MAP && ctx->pipe_mode && map_pipe_fast(b, o) >= 0) return grb_launch_pipe(ctx, b, ls, le, o, map_pipe_fast(b, o), valid_bits);
    if (MODE == M_MAP && b.uniform && o.k == 1) {
        switch (o.ops[0]) {
            case GRB_OP_MEAN:
                return launch_tile<MODE, GRB_OP_MEAN>(ctx, b, ls, le, o);
            case GRB_OP_SUM:
                return launch_tile<MODE, GRB_OP_SUM>(ctx, b, ls, le, o);
            case GRB_OP_SUM_NOT_EMPTY:
                return launch_tile<MODE, GRB_OP_SUM_NOT_EMPTY>(ctx, b, ls, le, o);
            case GRB_OP_COUNT:
                return launch_tile<MODE, GRB_OP_COUNT>(ctx, b, ls, le, o);
            default:
                break;
        }
    }
    return launch_tile<MODE, -1>(ctx, b, ls, le, o);
}

int enter(grb_ctx *ctx) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    return GRB_OK;
}

}  // namespace


extern "C" {

int grb_count_overlaps_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                             const uint32_t *ls, const uint32_t *le, uint32_t *counts) {
    GRB_TRY(enter(ctx));
    if (!left_offsets) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "count_overlaps: NULL left_offsets");
    Batch b;
    GRB_TRY(make_batch(ctx, idx, left_offsets, nseq, &b));
    if (b.total == 0) return GRB_OK;
    if (!ls || !le || !counts) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "count_overlaps: NULL buffer");
    InBuf dls, dle;
    OutBuf dc;
    GRB_TRY(dls.stage(ctx, ls, b.total * 4));
    GRB_TRY(dle.stage(ctx, le, b.total * 4));
    GRB_TRY(dc.stage(ctx, counts, b.total * 4));
    TileOut o;
    memset(&o, 0, sizeof(o));
    o.flags = ctx->dev_flags;
    o.counts = dc.as<u32>();
    GRB_TRY(launch_tile<M_COUNTS, -1>(ctx, b, dls.as<u32>(), dle.as<u32>(), o));
    GRB_TRY(dc.finish(ctx));
    if (dc.host) return grb_check_flags(ctx);
    return GRB_OK;
}

int grb_count_overlaps(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                       uint32_t *counts) {
    uint64_t off[2] = {0, nl};
    return grb_count_overlaps_batch(ctx, &idx, off, 1, ls, le, counts);
}

int grb_filter_overlaps_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                              const uint32_t *ls, const uint32_t *le, int anti, uint8_t *keep, uint64_t *n_kept) {
    GRB_TRY(enter(ctx));
    if (!left_offsets) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "filter_overlaps: NULL left_offsets");
    Batch b;
    GRB_TRY(make_batch(ctx, idx, left_offsets, nseq, &b));
    if (n_kept) *n_kept = 0;
    if (b.total == 0) return GRB_OK;
    if (!ls || !le || !keep) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "filter_overlaps: NULL buffer");
    InBuf dls, dle;
    OutBuf dk;
    TempBuf nk;
    GRB_TRY(dls.stage(ctx, ls, b.total * 4));
    GRB_TRY(dle.stage(ctx, le, b.total * 4));
    GRB_TRY(dk.stage(ctx, keep, b.total));
    TileOut o;
    memset(&o, 0, sizeof(o));
    o.flags = ctx->dev_flags;
    o.keep = dk.as<u8>();
    o.anti = anti;
    if (n_kept) {
        GRB_TRY(nk.alloc(ctx, 8));
        GRB_CUDA(ctx, cudaMemsetAsync(nk.p, 0, 8, ctx->stream));
        o.n_kept = nk.as<unsigned long long>();
    }
    GRB_TRY(launch_tile<M_KEEP, -1>(ctx, b, dls.as<u32>(), dle.as<u32>(), o));
    GRB_TRY(dk.finish(ctx));
    if (n_kept) GRB_CUDA(ctx, cudaMemcpyAsync(n_kept, nk.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (dk.host || n_kept) return grb_check_flags(ctx);
    return GRB_OK;
}

int grb_filter_overlaps_select(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                               const uint32_t *ls, const uint32_t *le, int anti, uint64_t *kept_index, uint64_t *n_kept) {
    GRB_TRY(enter(ctx));
    if (!left_offsets || !n_kept) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "filter_overlaps_select: NULL left_offsets / n_kept");
    const u64 total = left_offsets[nseq > 0 ? nseq : 0];
    *n_kept = 0;
    if (total == 0) return GRB_OK;
    if (!ls || !le || !kept_index) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "filter_overlaps_select: NULL buffer");
    InBuf dls, dle;
    GRB_TRY(dls.stage(ctx, ls, total * 4));
    GRB_TRY(dle.stage(ctx, le, total * 4));
    TempBuf keep, flags, pos;
    GRB_TRY(keep.alloc(ctx, total));
    GRB_TRY(flags.alloc(ctx, total * 4));
    GRB_TRY(pos.alloc(ctx, (total + 1) * 8));
    GRB_TRY(grb_filter_overlaps_batch(ctx, idx, left_offsets, nseq, dls.as<u32>(), dle.as<u32>(), anti, keep.as<u8>(), nullptr));
    const unsigned blocks = (unsigned)((total + 255) / 256);
    k_keep_to_u32<<<blocks, 256, 0, ctx->stream>>>(keep.as<u8>(), total, flags.as<u32>());
    GRB_LAUNCHED(ctx);
    GRB_TRY(grb_scan_u32_to_u64(ctx, flags.as<u32>(), pos.as<u64>(), total, true));
    u64 nk = 0;
    GRB_CUDA(ctx, cudaMemcpyAsync(&nk, pos.as<u64>() + total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    OutBuf dk;
    GRB_TRY(dk.stage(ctx, kept_index, nk * 8));
    if (nk) {
        k_scatter_kept<<<blocks, 256, 0, ctx->stream>>>(keep.as<u8>(), pos.as<u64>(), total, dk.as<u64>());
        GRB_LAUNCHED(ctx);
    }
    GRB_TRY(dk.finish(ctx));
    *n_kept = nk;
    return grb_check_flags(ctx);
}

int grb_filter_overlaps(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl, int anti,
                        uint8_t *keep, uint64_t *n_kept) {
    uint64_t off[2] = {0, nl};
    return grb_filter_overlaps_batch(ctx, &idx, off, 1, ls, le, anti, keep, n_kept);
}

static int map_joins_impl(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq, const uint32_t *ls,
                          const uint32_t *le, const int *ops, int k, double *out, uint8_t *out_valid, bool allow_chunks,
                          uint32_t *valid_bits = nullptr, bool no_bucket = false);

// Host buffers, many seqnames: run the seqnames in chunks so that the upload of the next chunk's lefts, the kernels
// of the current one and the download of the previous chunk's results overlap (PCIe is full duplex).
static int map_joins_host_chunked(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                                  const uint32_t *ls, const uint32_t *le, const int *ops, int k, double *out, uint8_t *out_valid) {
    const u64 total = left_offsets[nseq];
    TempBuf d_ls, d_le, d_out, d_val;
    GRB_TRY(d_ls.alloc(ctx, total * 4));
    GRB_TRY(d_le.alloc(ctx, total * 4));
    GRB_TRY(d_out.alloc(ctx, total * (u64)k * 8));
    GRB_TRY(d_val.alloc(ctx, total * (u64)k));
    // chunks of whole seqnames, about an eighth of the lefts each
    std::vector<int> cuts{0};
    const u64 target = total / 8 + 1;
    for (int s = 1; s < nseq; s++)
        if (left_offsets[s] - left_offsets[cuts.back()] >= target) cuts.push_back(s);
    cuts.push_back(nseq);
    const int nchunks = (int)cuts.size() - 1;
    cudaStream_t h2d, d2h;
    GRB_CUDA(ctx, cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    GRB_CUDA(ctx, cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> up((size_t)nchunks), done((size_t)nchunks);
    cudaEvent_t ready;
    cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
    cudaEventRecord(ready, ctx->stream);  // the staging buffers exist from here on
    cudaStreamWaitEvent(h2d, ready, 0);
    for (int c = 0; c < nchunks; c++) {
        cudaEventCreateWithFlags(&up[c], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming);
        const u64 a = left_offsets[cuts[c]], b = left_offsets[cuts[c + 1]];
        if (b > a) {
            cudaMemcpyAsync(d_ls.as<u32>() + a, ls + a, (b - a) * 4, cudaMemcpyHostToDevice, h2d);
            cudaMemcpyAsync(d_le.as<u32>() + a, le + a, (b - a) * 4, cudaMemcpyHostToDevice, h2d);
        }
        cudaEventRecord(up[c], h2d);
    }
    int rc = GRB_OK;
    for (int c = 0; c < nchunks && rc == GRB_OK; c++) {
        const u64 a = left_offsets[cuts[c]], b = left_offsets[cuts[c + 1]];
        cudaStreamWaitEvent(ctx->stream, up[c], 0);
        // absolute offsets + whole-array pointers: tiles stay aligned to multiples of 1024 lefts
        rc = map_joins_impl(ctx, idx ? idx + cuts[c] : nullptr, left_offsets + cuts[c], cuts[c + 1] - cuts[c], d_ls.as<u32>(),
                            d_le.as<u32>(), ops, k, d_out.as<double>(), d_val.as<u8>(), false);
        cudaEventRecord(done[c], ctx->stream);
        cudaStreamWaitEvent(d2h, done[c], 0);
        if (b > a && rc == GRB_OK) {
            cudaMemcpyAsync(out + a * (u64)k, d_out.as<double>() + a * (u64)k, (b - a) * (u64)k * 8, cudaMemcpyDeviceToHost, d2h);
            cudaMemcpyAsync(out_valid + a * (u64)k, d_val.as<u8>() + a * (u64)k, (b - a) * (u64)k, cudaMemcpyDeviceToHost, d2h);
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(h2d), e2 = cudaStreamSynchronize(d2h), e3 = cudaStreamSynchronize(ctx->stream);
    for (int c = 0; c < nchunks; c++) {
        cudaEventDestroy(up[c]);
        cudaEventDestroy(done[c]);
    }
    cudaEventDestroy(ready);
    cudaStreamDestroy(h2d);
    cudaStreamDestroy(d2h);
    if (rc != GRB_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return grb_fail(ctx, GRB_ERR_CUDA, "map_joins: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
    return grb_check_flags(ctx);
}

int grb_map_joins_f64_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                            const uint32_t *ls, const uint32_t *le, const int *ops, int k, double *out,
                            uint8_t *out_valid) {
    return map_joins_impl(ctx, idx, left_offsets, nseq, ls, le, ops, k, out, out_valid, true);
}

int grb_map_joins_f64_batch_bits(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                                 const uint32_t *ls, const uint32_t *le, const int *ops, int k, double *out,
                                 uint32_t *out_valid_bits) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    if (!out_valid_bits) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_bits: NULL validity buffer");
    if (left_offsets && nseq > 0 && left_offsets[0] != 0)
        return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_bits: left_offsets[0] must be 0 (the validity words are addressed from row 0)");
    return map_joins_impl(ctx, idx, left_offsets, nseq, ls, le, ops, k, out, nullptr, false, out_valid_bits);
}

__global__ void k_store_u32(u32 *p, u32 v) { *p = v; }

// validity bytes -> bits: word w covers flattened entries [32 w, 32 w + 32) of the nl x k byte array
__global__ void k_pack_bits(const u8 *__restrict__ bytes, u64 n, u32 *__restrict__ bits) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 m = __ballot_sync(GRB_FULL, i < n && bytes[i] != 0);
    if (lane_id() == 0 && (i & ~31ull) < n) bits[i >> 5] = m;
}

// Should the lefts of this call go through the coordinate buckets (bucket.cu)?  See grb_left_order in the header.
static int decide_bucketing(grb_ctx *ctx, const u32 *ls, u64 first, u64 total, u32 ntiles, bool *bucket, bool *report) {
    *bucket = *report = false;
    const u64 n = total - first;
    if (n < (1u << 16) || n >= 0xffffffffull) return GRB_OK;  // small calls: the searches are cheap either way
    if (ctx->left_order == GRB_LEFT_ORDER_SORTED) return GRB_OK;
    if (ctx->left_order == GRB_LEFT_ORDER_UNSORTED) {
        *bucket = true;
        return GRB_OK;
    }
    if (!grb_is_device_ptr(ls)) {
        int d = 0;
        GRB_TRY(grb_left_disorder(ctx, ls, first, total, &d));
        *bucket = d > 100;  // more than ~10 % of the sampled neighbours out of order
        return GRB_OK;
    }
    if (!ctx->stats_host) return GRB_OK;
    // Device buffers: what have earlier calls on this very buffer reported?  Reports arrive when their kernels finish, which may
    // be several (asynchronous) calls later, so the verdict in force is kept until a report that belongs to this buffer -- its
    // generation is not older than the buffer's first call -- says otherwise.
    volatile u32 *st = ctx->stats_host;
    if (ctx->hist_ls == (const void *)ls && ctx->hist_first == first && ctx->hist_total == total) {
        const u32 gen_bucket = st[2], gen_plain = st[1];
        const bool have_bucket = (u32)(gen_bucket - ctx->hist_gen0) < 0x40000000u, have_plain = (u32)(gen_plain - ctx->hist_gen0) < 0x40000000u;
        if (have_bucket && (!have_plain || (u32)(gen_bucket - gen_plain) < 0x40000000u)) {
            ctx->hist_bucketed = (u64)st[0] * 1024 / n > 100;  // adjacent inversions seen by the newest bucket pass
        } else if (have_plain) {
            u64 unstaged = 0;
            for (int b = 0; b < 1000; b++) unstaged += st[16 + b];
            ctx->hist_bucketed = unstaged * 8 > ctx->hist_tiles;  // more than an eighth of the tiles fell off the staged path
        }
    } else {
        ctx->hist_ls = ls;
        ctx->hist_first = first;
        ctx->hist_total = total;
        ctx->hist_bucketed = 0;
        ctx->hist_gen0 = ctx->gen + 1;
    }
    ctx->hist_tiles = ntiles;
    ctx->gen++;
    *bucket = ctx->hist_bucketed != 0;
    *report = true;
    return GRB_OK;
}

static int map_joins_impl(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq, const uint32_t *ls,
                          const uint32_t *le, const int *ops, int k, double *out, uint8_t *out_valid, bool allow_chunks,
                          uint32_t *valid_bits, bool no_bucket) {
    GRB_TRY(enter(ctx));
    if (!left_offsets || !ops) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins: NULL left_offsets / ops");
    if (k < 1 || k > GRB_MAX_OPS) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins: k must be in [1, %d]", GRB_MAX_OPS);
    bool need_scores = false, need_sum = false, need_members = false, want_median = false, want_bp = false, want_wm = false;
    for (int c = 0; c < k; c++) {
        switch (ops[c]) {
            case GRB_OP_SUM:
            case GRB_OP_SUM_NOT_EMPTY:
            case GRB_OP_MEAN:
                need_scores = need_sum = true;
                break;
            case GRB_OP_MIN:
            case GRB_OP_MAX:
                need_scores = need_members = true;
                break;
            case GRB_OP_MEDIAN:
                need_scores = need_members = want_median = true;
                break;
            case GRB_OP_COUNT:
                break;
            case GRB_OP_OVERLAP_BP:
                need_members = want_bp = true;
                break;
            case GRB_OP_WEIGHTED_MEAN:
            case GRB_OP_WEIGHTED_SUM:
                need_scores = need_members = want_bp = want_wm = true;
                break;
            default:
                return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins: unknown operation code %d", ops[c]);
        }
    }
    for (int s = 0; s < nseq; s++) {
        const grb_index *ix = idx ? idx[s] : nullptr;
        if (!ix) continue;
        if (need_scores && !ix->has_scores)
            return grb_fail(ctx, GRB_ERR_NO_DATA, "map_joins: the right index of seqname %d has no data container (scores)", s);
        if (want_median && ix->has_nan)
            return grb_fail(ctx, GRB_ERR_NAN_SCORE, "map_joins: median over NaN scores (the reference panics in partial_cmp().unwrap())");
        if (need_sum && ix->n && ix->sum_mode == GRB_SUM_SWEEP) need_members = true;
    }
    // WEIGHTED_MEAN has a closed form (k_weighted_mean) when every index has exact sums with room for score * coordinate
    u32 wm_cols = 0, ws_cols = 0;
    if (want_wm && ctx->sweep_mode != 0) {
        bool ok = true;
        for (int s = 0; s < nseq && ok; s++) {
            const grb_index *ix = idx ? idx[s] : nullptr;
            if (!ix || !ix->n) continue;
            GRB_TRY(grb_index_prepare_wm(ctx, const_cast<grb_index *>(ix)));
            ok = ix->wm_ready;
        }
        if (ok) {
            for (int c = 0; c < k; c++)
                if (ops[c] == GRB_OP_WEIGHTED_MEAN)
                    wm_cols |= 1u << c;
                else if (ops[c] == GRB_OP_WEIGHTED_SUM)
                    ws_cols |= 1u << c;
            want_wm = false;
        }
    }
    // OVERLAP_BP has a closed form (k_overlap_bp) unless a swept weighted mean needs the members' widths anyway
    const bool bp_closed = want_bp && !want_wm && ctx->sweep_mode != 0;
    u32 bp_cols = 0;
    if (bp_closed) {
        bool others = false;
        for (int c = 0; c < k; c++) {
            if (ops[c] == GRB_OP_OVERLAP_BP)
                bp_cols |= 1u << c;
            else if (ops[c] == GRB_OP_MIN || ops[c] == GRB_OP_MAX || ops[c] == GRB_OP_MEDIAN)
                others = true;
        }
        want_bp = false;
        need_members = others;
        for (int s = 0; s < nseq; s++) {
            const grb_index *ix = idx ? idx[s] : nullptr;
            if (!ix || !ix->n) continue;
            if (need_sum && ix->sum_mode == GRB_SUM_SWEEP) need_members = true;
            GRB_TRY(grb_index_prepare_bp(ctx, const_cast<grb_index *>(ix)));
        }
    }
    // MIN / MAX / MEDIAN alone (beside reductions the prefix sums answer) take the staged rank sweep; its tables
    // are built the first time an index is asked for them
    bool rank_sweep = ctx->sweep_mode != 0 && need_members && !want_bp && k <= 8;
    for (int s = 0; s < nseq && rank_sweep; s++) {
        const grb_index *ix = idx ? idx[s] : nullptr;
        if (!ix || !ix->n) continue;
        if (ix->has_nonfinite || ix->n >= 0x7fffffffu || (need_sum && ix->sum_mode == GRB_SUM_SWEEP)) rank_sweep = false;
    }
    for (int s = 0; s < nseq && rank_sweep; s++) {
        const grb_index *ix = idx ? idx[s] : nullptr;
        if (ix && ix->n && !ix->members_ready) GRB_TRY(grb_index_prepare_members(ctx, const_cast<grb_index *>(ix)));
    }
    // MIN / MAX without a median beside them need no sweep at all: stabbing tables + range minimum (k_minmax)
    u32 min_cols = 0, max_cols = 0;
    {
        if (rank_sweep && !want_median && ctx->sweep_mode == 1 && ctx->minmax_mode != 0) {
            for (int c = 0; c < k; c++) {
                if (ops[c] == GRB_OP_MIN) min_cols |= 1u << c;
                if (ops[c] == GRB_OP_MAX) max_cols |= 1u << c;
            }
            for (int s = 0; s < nseq; s++) {
                const grb_index *ix = idx ? idx[s] : nullptr;
                if (ix && ix->n && !ix->minmax_ready) GRB_TRY(grb_index_prepare_minmax(ctx, const_cast<grb_index *>(ix)));
            }
        }
    }
    Batch b;
    GRB_TRY(make_batch(ctx, idx, left_offsets, nseq, &b));
    if (b.total == 0 || b.ntiles == 0) return GRB_OK;  // (a batch that starts past 0 may hold no lefts at all)
    if (!ls || !le || !out || (!out_valid && !valid_bits)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins: NULL buffer");
    bool bucket = false, report = false;
    if (!no_bucket) GRB_TRY(decide_bucketing(ctx, ls, left_offsets[0], b.total, b.ntiles, &bucket, &report));
    if (bucket) {
        // lefts in coordinate buckets -> the same call on the bucketed copies -> rows back to where their lefts came from
        const u64 first = left_offsets[0], n = b.total - first;
        InBuf dls, dle;
        OutBuf dout, dval, dbits;
        TempBuf ls_b, le_b, rank, out_b, val_b, val_tmp, inv;
        GRB_TRY(dls.stage(ctx, ls + first, n * 4));
        GRB_TRY(dle.stage(ctx, le + first, n * 4));
        GRB_TRY(ls_b.alloc(ctx, b.total * 4));
        GRB_TRY(le_b.alloc(ctx, b.total * 4));
        GRB_TRY(rank.alloc(ctx, b.total * 4));
        GRB_TRY(out_b.alloc(ctx, b.total * (u64)k * 8));
        GRB_TRY(val_b.alloc(ctx, b.total * (u64)k));
        GRB_TRY(inv.alloc(ctx, 4));
        std::vector<u64> extent((size_t)nseq);
        for (int s = 0; s < nseq; s++) {
            const grb_index *ix = idx ? idx[s] : nullptr;
            extent[s] = ix && ix->n ? ((u64)ix->lut_bmax << ix->lut_shift) : 0;
        }
        // (the staged copies start at the batch's first left: shift the pointers back so that absolute positions index them)
        GRB_TRY(grb_bucket_lefts(ctx, left_offsets, extent.data(), nseq, dls.as<u32>() - first, dle.as<u32>() - first, ls_b.as<u32>(), le_b.as<u32>(),
                                 rank.as<u32>(), inv.as<u32>()));
        if (report) {
            // inversions, then -- in stream order, so it lands second -- the generation of this report
            GRB_CUDA(ctx, cudaMemcpyAsync(ctx->stats_host, inv.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            GRB_CUDA(ctx, cudaMemsetAsync(inv.p, 0, 4, ctx->stream));
            k_store_u32<<<1, 1, 0, ctx->stream>>>(ctx->stats_dev + 2, ctx->gen);
            GRB_LAUNCHED(ctx);
        }
        GRB_TRY(map_joins_impl(ctx, idx, left_offsets, nseq, ls_b.as<u32>(), le_b.as<u32>(), ops, k, out_b.as<double>(), val_b.as<u8>(), false, nullptr,
                               true));
        GRB_TRY(dout.stage(ctx, out + first * (u64)k, n * (u64)k * 8));
        u8 *vdst;
        if (valid_bits) {
            GRB_TRY(dbits.stage(ctx, valid_bits, ((b.total * (u64)k + 31) / 32) * 4));
            GRB_TRY(val_tmp.alloc(ctx, b.total * (u64)k));
            vdst = val_tmp.as<u8>();
        } else {
            GRB_TRY(dval.stage(ctx, out_valid + first * (u64)k, n * (u64)k));
            vdst = dval.as<u8>() - first * (u64)k;
        }
        GRB_TRY(grb_unbucket_results(ctx, first, b.total, k, out_b.as<double>(), val_b.as<u8>(), rank.as<u32>(), dout.as<double>() - first * (u64)k, vdst));
        if (valid_bits) {
            const u64 nflat = b.total * (u64)k;
            k_pack_bits<<<(unsigned)((nflat + 255) / 256), 256, 0, ctx->stream>>>(val_tmp.as<u8>(), nflat, dbits.as<u32>());
            GRB_LAUNCHED(ctx);
        }
        GRB_TRY(dout.finish(ctx));
        GRB_TRY(dval.finish(ctx));
        GRB_TRY(dbits.finish(ctx));
        if (dout.host || dval.host || dbits.host) return grb_check_flags(ctx);
        return GRB_OK;
    }
    if (allow_chunks && nseq >= 2 && b.total >= (1u << 22) && !grb_is_device_ptr(ls) && !grb_is_device_ptr(le) &&
        !grb_is_device_ptr(out) && !grb_is_device_ptr(out_valid))
        return map_joins_host_chunked(ctx, idx, left_offsets, nseq, ls, le, ops, k, out, out_valid);
    InBuf dls, dle;
    OutBuf dout, dval, dbits;
    TempBuf val_tmp;
    GRB_TRY(dls.stage(ctx, ls, b.total * 4));
    GRB_TRY(dle.stage(ctx, le, b.total * 4));
    GRB_TRY(dout.stage(ctx, out, b.total * (u64)k * 8));
    TileOut o;
    memset(&o, 0, sizeof(o));
    o.flags = ctx->dev_flags;
    o.stats = report ? ctx->stats_dev : nullptr;
    o.stats_gen = ctx->gen;
    o.out = dout.as<double>();
    o.k = k;
    o.need_sum = need_sum;
    for (int c = 0; c < k; c++) o.ops[c] = ops[c];
    // validity as bits: written by the pipelined kernel itself for a single prefix-sum reduction, else packed from bytes
    bool bits_direct = false;
    if (valid_bits) {
        GRB_TRY(dbits.stage(ctx, valid_bits, ((b.total * (u64)k + 31) / 32) * 4));
        bits_direct = k == 1 && !need_members && !bp_cols && !wm_cols && !ws_cols && ctx->pipe_mode == 1 && map_pipe_fast(b, o) >= 0;
        if (!bits_direct) {
            GRB_TRY(val_tmp.alloc(ctx, b.total * (u64)k));
            o.out_valid = val_tmp.as<u8>();
        }
    } else {
        GRB_TRY(dval.stage(ctx, out_valid, b.total * (u64)k));
        o.out_valid = dval.as<u8>();
    }
    if (!need_members && !bp_cols && !wm_cols && !ws_cols) {
        GRB_TRY(launch_map<M_MAP>(ctx, b, dls.as<u32>(), dle.as<u32>(), o, bits_direct ? dbits.as<u32>() : nullptr));
    } else {
        TempBuf ub, lb, pp, heavy;
        GRB_TRY(ub.alloc(ctx, b.total * 4));
        GRB_TRY(lb.alloc(ctx, b.total * 4));
        GRB_TRY(pp.alloc(ctx, b.total * 4));
        o.ub = ub.as<u32>();
        o.lb = lb.as<u32>();
        o.pp = pp.as<u32>();
        // k_sweep3 follows and some columns come from the prefix sums: the pipelined pass leaves (sum, count, scored members)
        // per left in a side array and the sweep writes whole rows -- otherwise both kernels read-modify-write the partial
        // sectors of a row-major `out` (at config C3: 4.5 GB of extra DRAM reads in each of them)
        TempBuf planar;
        const bool ublb_in_pipe = ctx->pipe_mode == 1 && (ctx->ublb_pipe < 0 ? need_sum != 0 : ctx->ublb_pipe != 0) && (b.pipeable || !need_sum) &&
                                  b.ntiles >= 4u * (u32)ctx->sm_count;
        bool has_pipe_col = false;
        for (int c = 0; c < k; c++) has_pipe_col |= ops[c] == GRB_OP_COUNT || ops[c] == GRB_OP_SUM || ops[c] == GRB_OP_SUM_NOT_EMPTY || ops[c] == GRB_OP_MEAN;
        if (rank_sweep && want_median && ublb_in_pipe && has_pipe_col && !bp_cols && !wm_cols && !ws_cols && ctx->planar_rows) {
            GRB_TRY(planar.alloc(ctx, b.total * 16));
            o.planar = planar.as<uint4>();
        }
        GRB_TRY(launch_map<M_MAP | M_UBLB>(ctx, b, dls.as<u32>(), dle.as<u32>(), o));
        if (bp_cols) {
            k_overlap_bp<<<(unsigned)((b.total - left_offsets[0] + 255) / 256), 256, 0, ctx->stream>>>(b.seqs, b.nseq, left_offsets[0], b.total,
                                                                                                     dls.as<u32>(), dle.as<u32>(), o.ub,
                                                                                   o.pp, o.lb, o.out, o.out_valid, k, bp_cols);
            GRB_LAUNCHED(ctx);
        }
        if (wm_cols | ws_cols) {
            k_weighted_mean<<<(unsigned)((b.total - left_offsets[0] + 255) / 256), 256, 0, ctx->stream>>>(b.seqs, b.nseq, left_offsets[0], b.total,
                                                                                                        dls.as<u32>(), dle.as<u32>(),
                                                                                      o.ub, o.pp, o.lb, o.out, o.out_valid, k, wm_cols, ws_cols);
            GRB_LAUNCHED(ctx);
        }
        if (min_cols | max_cols) {
            k_minmax<<<(unsigned)((b.total - left_offsets[0] + 255) / 256), 256, 0, ctx->stream>>>(b.seqs, b.nseq, left_offsets[0], b.total, o.ub, o.pp,
                                                                                                 o.lb, o.out, o.out_valid, k, min_cols, max_cols);
            GRB_LAUNCHED(ctx);
        } else if (need_members) {
            SweepOut so;
            memset(&so, 0, sizeof(so));
            so.out = o.out;
            so.out_valid = o.out_valid;
            so.k = k;
            for (int c = 0; c < k; c++) so.ops[c] = ops[c];
            so.want_median = want_median;
            so.want_bp = want_bp;
            so.want_wm = want_wm;
            so.wave_ratio = ctx->wave_ratio;
            so.planar = o.planar;
            if (want_median) {
                if (b.total >= 0xffffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "map_joins: median over more than 2^32-1 lefts per call");
                GRB_TRY(heavy.alloc(ctx, (b.total + 1) * 4));
                GRB_CUDA(ctx, cudaMemsetAsync(heavy.p, 0, 4, ctx->stream));
                so.heavy_count = heavy.as<u32>();
                so.heavy_list = heavy.as<u32>() + 1;
            }
            if (rank_sweep) {
                static bool configured[64] = {};
                if (ctx->device < 0 || ctx->device >= 64 || !configured[ctx->device]) {
                    GRB_CUDA(ctx, cudaFuncSetAttribute(k_sweep3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sweep3Smem)));
                    if (ctx->device >= 0 && ctx->device < 64) configured[ctx->device] = true;
                }
                k_sweep3<<<b.ntiles, SW3_TPB, sizeof(Sweep3Smem), ctx->stream>>>(b.seqs, b.tile_begin, b.nseq, dls.as<u32>(), o.ub, o.pp, o.lb, so,
                                                                                ctx->sweep_mode == 1 ? 1 : 0);
            } else {
                k_sweep2<<<b.ntiles, SW2_TPB, 0, ctx->stream>>>(b.seqs, b.tile_begin, b.nseq, dls.as<u32>(), dle.as<u32>(), o.ub, o.pp, o.lb, so);
            }
            GRB_LAUNCHED(ctx);
            if (want_median) {
                if (rank_sweep)
                    k_median_heavy3<<<ctx->sm_count * 6, MH3_TPB, 0, ctx->stream>>>(b.seqs, b.nseq, dls.as<u32>(), o.ub, o.pp, o.lb,
                                                                                    so.heavy_list, so.heavy_count, so);
                else
                    k_median_heavy<<<ctx->sm_count * 4, SW_TPB, 0, ctx->stream>>>(b.seqs, b.nseq, dls.as<u32>(), o.ub, o.lb, so.heavy_list,
                                                                                  so.heavy_count, so);
                GRB_LAUNCHED(ctx);
            }
        }
    }
    if (valid_bits && !bits_direct) {
        const u64 nflat = b.total * (u64)k;
        k_pack_bits<<<(unsigned)((nflat + 255) / 256), 256, 0, ctx->stream>>>(val_tmp.as<u8>(), nflat, dbits.as<u32>());
        GRB_LAUNCHED(ctx);
    }
    GRB_TRY(dout.finish(ctx));
    GRB_TRY(dval.finish(ctx));
    GRB_TRY(dbits.finish(ctx));
    if (dout.host || dval.host || dbits.host) return grb_check_flags(ctx);
    return GRB_OK;
}

int grb_map_joins_f64(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                      const int *ops, int k, double *out, uint8_t *out_valid) {
    uint64_t off[2] = {0, nl};
    return grb_map_joins_f64_batch(ctx, &idx, off, 1, ls, le, ops, k, out, out_valid);
}

int grb_left_overlaps(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                      uint64_t *offsets, grb_pairs **out) {
    GRB_TRY(enter(ctx));
    if (!out || !offsets) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "left_overlaps: NULL output");
    *out = nullptr;
    if (nl && (!ls || !le)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "left_overlaps: NULL buffer");
    grb_pairs *p = (grb_pairs *)calloc(1, sizeof(grb_pairs));
    if (!p) return grb_fail(ctx, GRB_ERR_OOM, "left_overlaps: host allocation failed");
    p->ctx = ctx;
    p->idx = idx;
    p->nl = nl;
    int rc = GRB_OK;
    do {
        if (grb_dev_malloc(ctx, (void **)&p->ls, nl * 4 + 16) != cudaSuccess || grb_dev_malloc(ctx, (void **)&p->le, nl * 4 + 16) != cudaSuccess ||
            grb_dev_malloc(ctx, (void **)&p->offsets, (nl + 1) * 8) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_OOM, "left_overlaps: device allocation failed");
            break;
        }
        if (nl) {
            cudaMemcpyAsync(p->ls, ls, nl * 4, cudaMemcpyDefault, ctx->stream);
            cudaMemcpyAsync(p->le, le, nl * 4, cudaMemcpyDefault, ctx->stream);
        }
        uint64_t off[2] = {0, nl};
        Batch b;
        if ((rc = make_batch(ctx, &idx, off, 1, &b))) break;
        TempBuf counts, ub, lb, pp;
        if ((rc = counts.alloc(ctx, nl * 4))) break;
        if ((rc = ub.alloc(ctx, nl * 4))) break;
        if ((rc = lb.alloc(ctx, nl * 4))) break;
        if ((rc = pp.alloc(ctx, nl * 4))) break;
        TileOut o;
        memset(&o, 0, sizeof(o));
        o.flags = ctx->dev_flags;
    o.flags = ctx->dev_flags;
        o.counts = counts.as<u32>();
        o.ub = ub.as<u32>();
        o.lb = lb.as<u32>();
        o.pp = pp.as<u32>();
        if ((rc = launch_tile<M_COUNTS | M_UBLB, -1>(ctx, b, p->ls, p->le, o))) break;
        if ((rc = grb_scan_u32_to_u64(ctx, counts.as<u32>(), p->offsets, nl, true))) break;
        u64 np = 0;
        cudaMemcpyAsync(&np, p->offsets + nl, 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "left_overlaps: %s", cudaGetErrorString(e));
            break;
        }
        if ((rc = grb_check_flags(ctx))) break;  // an invalid left range: report before materialising anything
        p->np = np;
        if (grb_dev_malloc(ctx, (void **)&p->pos, np * 4 + 16) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_OOM, "left_overlaps: cannot allocate %llu pairs on the device", (unsigned long long)np);
            break;
        }
        if (np) {
            k_emit2<<<b.ntiles, SW2_TPB, 0, ctx->stream>>>(b.seqs, b.tile_begin, 1, p->ls, o.ub, o.pp, o.lb, p->offsets, p->pos);
            GRB_LAUNCHED_BREAK(ctx, rc)
        }
        e = cudaMemcpyAsync(offsets, p->offsets, (nl + 1) * 8, cudaMemcpyDefault, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "left_overlaps: %s", cudaGetErrorString(e));
            break;
        }
    } while (0);
    if (rc != GRB_OK) {
        grb_pairs_destroy(p);
        return rc;
    }
    *out = p;
    return GRB_OK;
}

size_t grb_pairs_len(const grb_pairs *p) { return p ? (size_t)p->np : 0; }

int grb_pairs_copy(const grb_pairs *p, uint64_t *right_data_idx, uint32_t *r_start, uint32_t *r_end) {
    if (!p) return GRB_ERR_INVALID_ARG;
    grb_ctx *ctx = p->ctx;
    GRB_TRY(enter(ctx));
    if (p->np == 0) return GRB_OK;
    OutBuf di, ds, de;
    GRB_TRY(di.stage(ctx, right_data_idx, p->np * 8));
    GRB_TRY(ds.stage(ctx, r_start, p->np * 4));
    GRB_TRY(de.stage(ctx, r_end, p->np * 4));
    u64 blocks = (p->np + 255) / 256;
    k_pairs_gather<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p->pos, p->np, p->idx->starts, p->idx->ends, p->idx->didx, di.as<u64>(),
                                                             ds.as<u32>(), de.as<u32>());
    GRB_LAUNCHED(ctx);
    GRB_TRY(di.finish(ctx));
    GRB_TRY(ds.finish(ctx));
    GRB_TRY(de.finish(ctx));
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

int grb_pairs_overlap_widths(const grb_pairs *p, uint32_t *widths) {
    if (!p || !widths) return GRB_ERR_INVALID_ARG;
    grb_ctx *ctx = p->ctx;
    GRB_TRY(enter(ctx));
    if (p->np == 0) return GRB_OK;
    OutBuf dw;
    GRB_TRY(dw.stage(ctx, widths, p->np * 4));
    u64 blocks = (p->nl * 32 + 255) / 256;
    k_pairs_widths<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p->pos, p->offsets, p->nl, p->ls, p->le, p->idx->starts, p->idx->ends,
                                                             dw.as<u32>());
    GRB_LAUNCHED(ctx);
    GRB_TRY(dw.finish(ctx));
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

void grb_pairs_destroy(grb_pairs *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    grb_dev_free(p->ctx, p->ls);
    grb_dev_free(p->ctx, p->le);
    grb_dev_free(p->ctx, p->offsets);
    grb_dev_free(p->ctx, p->pos);
    free(p);
}

}  // extern "C"
