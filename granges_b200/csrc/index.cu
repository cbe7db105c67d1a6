// index.cu -- the right-hand index that replaces COITrees<M>.
//
// Reference: COITrees::from(VecRanges<R>) (src/ranges/coitrees.rs:78-84), reached from
// GRanges::into_coitrees (src/granges.rs:1365-1402); closed/half-open conversion
// src/ranges/coitrees.rs:172-198; the map_data score extraction src/commands.rs:514-517.
//
// Layout in HBM, per seqname (n rights):
//   start order  starts[n] ends[n] didx[n]      sorted by (start, end, data index) = sort_ranges order
//                bme[n/64]                      block-max-end (replaces coitrees' subtree_last)
//   end order    ends_sorted[n] perm_e[n]       the ends sorted on their own
//   scores       score_a[n] (+ validity bits and prefix counts when some scores are None)
//   exact sums   fixed-point prefix sums of the scores in BOTH orders, so that
//                sum(overlapping) = PA[#{r.start < l.end}] - PB[#{r.end <= l.start}]
//                compact: i64 prefix local to a block of 256 + one i128 base per block
//                wide:    one i128 prefix per right
// count_overlaps is the same identity on the two key arrays:
//   #{r.start < l.end} - #{r.end <= l.start}   (exact: r.end <= l.start implies r.start < l.end).
#include <math.h>
#include <stdlib.h>

#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr u32 PAD = 16;  // sentinel entries after the last key (bulk copies round up to 16 B)

struct BuildFlags {
    u32 invalid_range;
    u32 idx_too_big;
    u32 max_end;    // largest end seen (k_make_keys)
    u32 heavy_bin;  // the counting sort of the ends met a coordinate bin too crowded for its fix-up pass
};

constexpr u32 BIN_FIX_MAX = 128;  // ends per coordinate bin the placement pass of the counting sort orders by counting

__global__ void __launch_bounds__(256) k_make_keys(const u32 *__restrict__ starts, const u32 *__restrict__ ends, const u32 *__restrict__ order,
                                                   size_t n, u64 *__restrict__ keys, u32 *__restrict__ vals, BuildFlags *flags) {
    // grid-stride: a few blocks per SM, so that the largest end costs one atomic per block (every warp of a start-sorted
    // input raises the running maximum: per-warp atomics on the one word serialise -- 74 us of a 4.5M-range build)
    __shared__ u32 s_max[8];
    u32 mx = 0;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u32 src = order ? order[i] : (u32)i;
        const u32 s = starts[src], e = ends[src];
        bad |= e <= s || e > 0x7fffffffu;
        keys[i] = ((u64)s << 32) | e;
        vals[i] = src;
        mx = max(mx, e);
    }
    mx = warp_max(mx);
    if (__any_sync(GRB_FULL, bad) && lane_id() == 0) flags->invalid_range = 1;
    if (lane_id() == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) mx = max(mx, s_max[w]);
        atomicMax(&flags->max_end, mx);
    }
}

// ---- end order by a counting sort over the coordinate bins of the lookup table ---------------------------------------
// The table needs lut_b[b] = #ends < (b << shift) anyway: that IS the exclusive scan of the per-bin counts, and it also
// tells every end where its bin starts in end order.  So: count per bin (atomics; for start-sorted input neighbouring
// threads hit neighbouring bins, which the L2 coalesces), scan, scatter through per-bin cursors into scratch arrays, then
// place every end inside its bin -- ~2 ends on average -- by (end, position): exactly what a stable radix sort of (end, position) gives, in 4 light
// passes instead of 4 x (histogram + scan + scatter).  A bin beyond BIN_FIX_MAX ends (thousands of identical ends) raises a
// flag and the radix sort redoes the order.
__global__ void k_bin_count(const u32 *__restrict__ ends, u32 n, int shift, u32 bmax, u32 *__restrict__ cnt) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[min(ends[i] >> shift, bmax)], 1u);
}
__global__ void k_bin_scatter(const u32 *__restrict__ ends, u32 n, int shift, u32 bmax, const u32 *__restrict__ lut, u32 *__restrict__ cur,
                              u32 *__restrict__ ends_sorted, u32 *__restrict__ perm_e) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 e = ends[i];
    const u32 b = min(e >> shift, bmax);
    const u32 p = lut[b] + atomicAdd(&cur[b], 1u);
    ends_sorted[p] = e;
    perm_e[p] = i;
}
// Order inside the bins, one thread per end: the scatter leaves the bins' ends in arrival order in two scratch arrays; every
// end counts the members of its bin that precede it in (end, position) order and goes to that place of the
// final arrays.  Regular work (bins hold ~2 ends), coalesced reads and near-coalesced writes (one thread per bin sorting it in
// place took 72 us per 4.5M rights).
__global__ void k_bin_place(const u32 *__restrict__ lut, int shift, u32 bmax, const u32 *__restrict__ tmp_e, const u32 *__restrict__ tmp_p, u32 n,
                            u32 *__restrict__ ends_sorted, u32 *__restrict__ perm_e, BuildFlags *flags) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 e = tmp_e[i], q = tmp_p[i];
    const u32 b = min(e >> shift, bmax);
    const u32 lo = lut[b], hi = lut[b + 1];
    u32 dest = i;
    if (hi - lo > BIN_FIX_MAX) {
        flags->heavy_bin = 1;  // (the caller sorts the ends again with the radix sort)
    } else {
        u32 rank = 0;
        for (u32 j = lo; j < hi; j++) {
            const u32 ej = tmp_e[j], qj = tmp_p[j];
            rank += (ej < e || (ej == e && qj < q)) ? 1u : 0u;
        }
        dest = lo + rank;
    }
    ends_sorted[dest] = e;
    perm_e[dest] = q;
}
__global__ void k_fill_u32(u32 *__restrict__ p, u32 n, u32 v) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_idx_keys(const u64 *__restrict__ data_idx, size_t n, u64 *__restrict__ keys, u32 *__restrict__ vals,
                           BuildFlags *flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 d = data_idx[i];
    if (d >= 0xffffffffull) flags->idx_too_big = 1;
    keys[i] = d;
    vals[i] = (u32)i;
}

__global__ void k_unpack_start_order(const u64 *__restrict__ keys, const u32 *__restrict__ perm,
                                     const u64 *__restrict__ data_idx, size_t n, u32 *__restrict__ starts,
                                     u32 *__restrict__ ends, u32 *__restrict__ didx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + PAD) return;
    if (i >= n) {
        starts[i] = 0xffffffffu;
        ends[i] = 0xffffffffu;
        return;
    }
    u64 k = keys[i];
    u32 p = perm[i];
    starts[i] = (u32)(k >> 32);
    ends[i] = (u32)k;
    didx[i] = data_idx ? (u32)data_idx[p] : p;
}

__global__ void k_pad_keys(u32 *__restrict__ tail) { tail[threadIdx.x] = 0xffffffffu; }

__global__ void k_block_max_end(const u32 *__restrict__ ends, size_t n, u32 *__restrict__ bme) {
    // one warp per block of 64 rights
    size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nb = (n + 63) >> GRB_BME_SHIFT;
    if (b >= nb) return;
    size_t i0 = b << GRB_BME_SHIFT;
    u32 m = 0;
    for (int k = 0; k < 2; k++) {
        size_t i = i0 + k * 32 + lane_id();
        if (i < n) m = max(m, ends[i]);
    }
    m = warp_max(m);
    if (lane_id() == 0) bme[b] = m;
}

// lut[b] = #keys < (b << shift) for b in [0, bmax]; entries bmax+1 .. bmax+PAD = n.
// Position i is the answer for every bin b with bin(keys[i-1]) < b <= bin(keys[i]) (keys sorted; bin(keys[-1]) = -1,
// and the n-th "key" closes the table): each thread fills the bins between its key and its predecessor's -- coalesced
// reads, no searching.  The table has about n / 2 bins, so the fills are short on average.
__global__ void k_build_lut(const u32 *__restrict__ keys, u32 n, int shift, u32 bmax, u32 *__restrict__ lut, u16 *__restrict__ lut16) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const long long prev = i == 0 ? -1ll : (long long)min(keys[i - 1] >> shift, bmax);
    const long long cur = i == n ? (long long)bmax + PAD : (long long)min(keys[i] >> shift, bmax);
    for (long long b = prev + 1; b <= cur; b++) {
        lut[b] = i;
        lut16[b] = (u16)i;
    }
}

// ---- scores ------------------------------------------------------------------

struct ScoreStats {
    int min_lsb_exp;  // min over finite non-zero scores of the exponent of the lowest set bit
    int max_exp;      // max exponent of the highest set bit
    u32 n_null;
    u32 has_nonfinite;
    u32 has_nan;
    u32 has_subnormal;
    u32 idx_out_of_range;
    u32 n_nonzero;
};

// Four elements per thread (element (4 blockIdx + k) 256 + threadIdx: a warp still owns 32 consecutive positions per k, i.e. one
// validity word): the index loads, then the score gathers, of all four are in flight together.
__global__ void __launch_bounds__(256) k_gather_scores(const u32 *__restrict__ didx, size_t n, const double *__restrict__ scores,
                                                       const u8 *__restrict__ valid, u64 n_data, double *__restrict__ score_a,
                                                       u32 *__restrict__ vbits_a, u32 *__restrict__ vflag_a, ScoreStats *st) {
    int lsb = 1 << 30, top = -(1 << 30);
    u32 flags = 0, nnull = 0;
    size_t idx[4];
    u32 d[4];
    bool in[4], ok[4];
    double x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        idx[k] = ((size_t)blockIdx.x * 4 + k) * 256 + threadIdx.x;
        in[k] = idx[k] < n;
        d[k] = in[k] ? didx[idx[k]] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        ok[k] = false;
        x[k] = 0.0;
        if (in[k]) {
            if ((u64)d[k] >= n_data) {
                flags |= 8u;
            } else {
                ok[k] = valid ? valid[d[k]] != 0 : true;
                if (ok[k]) x[k] = scores[d[k]];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (in[k]) score_a[idx[k]] = x[k];
        if (ok[k]) {
            const u64 b = (u64)__double_as_longlong(x[k]);
            const int ex = (int)((b >> 52) & 0x7ff);
            const u64 frac = b & ((1ull << 52) - 1);
            if (ex == 0x7ff) {
                flags |= frac ? 3u : 1u;  // non-finite (and NaN)
            } else if (ex == 0) {
                if (frac) flags |= 4u;    // subnormal
            } else {
                const u64 m = frac | (1ull << 52);
                const int tz = __ffsll((long long)m) - 1;
                lsb = min(lsb, ex - 1075 + tz);
                top = max(top, ex - 1023);
                flags |= 16u;
            }
        }
        const u32 okmask = __ballot_sync(GRB_FULL, ok[k]);
        if (in[k]) {
            if (vbits_a && lane_id() == 0) vbits_a[idx[k] >> 5] = okmask;
            if (vflag_a) vflag_a[idx[k]] = ok[k] ? 1u : 0u;
        }
        nnull += __popc(__ballot_sync(GRB_FULL, in[k] && !ok[k]));
    }
    // warp reduce stats (REDUX: three instructions per warp instead of seven shuffle trees)
    lsb = __reduce_min_sync(GRB_FULL, lsb);
    top = __reduce_max_sync(GRB_FULL, top);
    const u32 wflags = __reduce_or_sync(GRB_FULL, flags);
    // block reduce, then one set of atomics per block
    __shared__ int s_lsb[8], s_top[8];
    __shared__ u32 s_flags[8], s_null[8];
    const u32 w = threadIdx.x >> 5;
    if (lane_id() == 0) {
        s_lsb[w] = lsb;
        s_top[w] = top;
        s_flags[w] = wflags;
        s_null[w] = nnull;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 fl = 0;
        for (u32 k = 0; k < (blockDim.x >> 5); k++) {
            lsb = min(lsb, s_lsb[k]);
            top = max(top, s_top[k]);
            fl |= s_flags[k];
            if (k) nnull += s_null[k];
        }
        volatile ScoreStats *vs = st;
        if (fl & 16u) {
            if (lsb < vs->min_lsb_exp) atomicMin(&st->min_lsb_exp, lsb);
            if (top > vs->max_exp) atomicMax(&st->max_exp, top);
            if (!vs->n_nonzero) atomicOr(&st->n_nonzero, 1u);
        }
        if (nnull) atomicAdd(&st->n_null, nnull);
        if ((fl & 1u) && !vs->has_nonfinite) atomicOr(&st->has_nonfinite, 1u);
        if ((fl & 2u) && !vs->has_nan) atomicOr(&st->has_nan, 1u);
        if ((fl & 4u) && !vs->has_subnormal) atomicOr(&st->has_subnormal, 1u);
        if ((fl & 8u) && !vs->idx_out_of_range) atomicOr(&st->idx_out_of_range, 1u);
    }
}

__global__ void k_permute_u32(const u32 *__restrict__ src, const u32 *__restrict__ perm, size_t n, u32 *__restrict__ dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

// Block-of-256 fixed-point prefix.  Positions p in [b*256, b*256+256) with p <= n; the value at
// position p is the sum of the block's elements before p.
//   phase 0: blocksum[b] = sum of the block's elements
//   phase 1 (compact): pfx[p] = low 64 bits of base[b] + local exclusive prefix
//   phase 2 (wide):    wide[p] = base[b] + local exclusive prefix (i128)
//   phase 3 (wide96):  pfx[p] / pfh[p] = bits 0..63 / 64..95 of base[b] + local exclusive prefix
// Both orders in one launch: blocks [0, nb) take the start order (side a, no permutation), blocks [nb, 2 nb) the end
// order (side b, through perm_e) -- half the launches of a build's prefix passes, each twice as large.
struct FxSide {
    const u32 *perm;
    i128 *blocksum;
    const i128 *base;
    u64 *pfx;
    i128 *wide;
    u32 *pfh;
};
// One WARP per block of 256 rights: lane l holds elements 8 l .. 8 l + 7 (64 contiguous bytes of scores and of prefixes per
// lane), scans them in registers and joins one 128-bit warp scan -- an eighth of the shuffles of one element per thread.
template <int PHASE>
__global__ void __launch_bounds__(256) k_fx_blocks(const double *__restrict__ score_a, size_t n, int e0, size_t nb, const FxSide sa,
                                                   const FxSide sb) {
    const size_t wid = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);  // warps [0, nb): side a, [nb, 2 nb): side b
    if (wid >= 2 * nb) return;
    const bool second = wid >= nb;
    const size_t b = second ? wid - nb : wid;
    const u32 *__restrict__ perm = second ? sb.perm : sa.perm;
    i128 *__restrict__ blocksum = second ? sb.blocksum : sa.blocksum;
    const i128 *__restrict__ base = second ? sb.base : sa.base;
    u64 *__restrict__ pfx = second ? sb.pfx : sa.pfx;
    i128 *__restrict__ wide = second ? sb.wide : sa.wide;
    u32 *__restrict__ pfh = second ? sb.pfh : sa.pfh;
    const u32 lane = lane_id();
    const size_t p0 = (b << GRB_BASE_SHIFT) + (size_t)lane * 8;
    // the lane's eight values, then their inclusive scan in place
    i128 v[8];
    if (p0 + 8 <= n) {
        double x[8];
        if (perm) {
            const uint4 q0 = *reinterpret_cast<const uint4 *>(perm + p0), q1 = *reinterpret_cast<const uint4 *>(perm + p0 + 4);
            const u32 idx[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int r = 0; r < 8; r++) x[r] = score_a[idx[r]];
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double2 d = *reinterpret_cast<const double2 *>(score_a + p0 + 2 * r);
                x[2 * r] = d.x;
                x[2 * r + 1] = d.y;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = to_fx(x[r], e0);
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const size_t p = p0 + r;
            v[r] = p < n ? to_fx(score_a[perm ? perm[p] : p], e0) : (i128)0;
        }
    }
    i128 ex[8];  // exclusive prefixes inside the lane
    i128 tot = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        ex[r] = tot;
        tot += v[r];
    }
    // warp inclusive scan of the lane totals through their two halves
    i128 incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 tlo = __shfl_up_sync(GRB_FULL, (u64)incl, o);
        const u64 thi = __shfl_up_sync(GRB_FULL, (u64)((u128)incl >> 64), o);
        if (lane >= (u32)o) incl += (i128)(((u128)thi << 64) | tlo);
    }
    if (PHASE == 0) {
        if (lane == 31) blocksum[b] = incl;
        return;
    }
    const i128 off = base[b] + (incl - tot);
    if (PHASE == 1) {
        if (p0 + 8 <= n + 1) {
#pragma unroll
            for (int r = 0; r < 4; r++)
                *reinterpret_cast<ulonglong2 *>(pfx + p0 + 2 * r) = make_ulonglong2((u64)(u128)(off + ex[2 * r]), (u64)(u128)(off + ex[2 * r + 1]));
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++)
                if (p0 + r <= n) pfx[p0 + r] = (u64)(u128)(off + ex[r]);
        }
    } else if (PHASE == 3) {
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (p0 + r <= n) {
                const u128 t = (u128)(off + ex[r]);
                pfx[p0 + r] = (u64)t;
                pfh[p0 + r] = (u32)(t >> 64);
            }
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (p0 + r <= n) wide[p0 + r] = off + ex[r];
    }
}

// out[i] = i: the valid-score counts of an index without None scores (only built beside non-finite scores, whose
// seqnames run in the NULLS variant of the pipelined kernel)
__global__ void k_iota_u32(size_t n, u32 *__restrict__ a, u32 *__restrict__ b) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = b[i] = (u32)i;
}

// low halves of a u32 array (what the pipelined kernel stages for the valid-score counts)
__global__ void k_narrow_u16(const u32 *__restrict__ in, size_t n, u16 *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u16)in[i];
}

// which = 0 / 1 / 2: flag the +inf / -inf / NaN scores (position i of the start order, or of the end order through perm)
__global__ void k_nf_flags(const double *__restrict__ score_a, const u32 *__restrict__ vbits_a, const u32 *__restrict__ perm, size_t n,
                           int which, u32 *__restrict__ flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 j = perm ? perm[i] : (u32)i;
    const bool ok = vbits_a ? ((vbits_a[j >> 5] >> (j & 31)) & 1u) : true;
    const u64 b = (u64)__double_as_longlong(score_a[j]);
    const bool nonfinite = ((b >> 52) & 0x7ff) == 0x7ff, nan = nonfinite && (b & ((1ull << 52) - 1)) != 0;
    const int cls = !ok || !nonfinite ? -1 : nan ? 2 : (b >> 63) ? 1 : 0;
    flags[i] = cls == which ? 1u : 0u;
}
__global__ void k_nf_pack(const u32 *__restrict__ a, const u32 *__restrict__ b, const u32 *__restrict__ c, size_t n, uint4 *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_uint4(a[i], b[i], c[i], 0u);
}

// The same block scan over score * coordinate (exact: the product of a W-bit and a 31-bit integer).
//   phase 0: blocksum[b];  phase 2: wide[p] = base[b] + local exclusive prefix, p in [b*256, b*256+256), p <= n
template <int PHASE>
__global__ void __launch_bounds__(256) k_wfx_blocks(const double *__restrict__ score_a, const u32 *__restrict__ perm,
                                                    const u32 *__restrict__ coord, size_t n, int e0, i128 *__restrict__ blocksum,
                                                    const i128 *__restrict__ base, i128 *__restrict__ wide) {
    __shared__ u64 s_lo[8], s_hi[8];
    const size_t b = blockIdx.x;
    const size_t p = (b << GRB_BASE_SHIFT) + threadIdx.x;
    i128 v = 0;
    if (p < n) v = to_fx(score_a[perm ? perm[p] : p], e0) * (i128)coord[p];
    i128 incl = v;
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 tlo = __shfl_up_sync(GRB_FULL, (u64)incl, o);
        u64 thi = __shfl_up_sync(GRB_FULL, (u64)((u128)incl >> 64), o);
        if (lane >= o) incl += (i128)(((u128)thi << 64) | tlo);
    }
    if (lane == 31) {
        s_lo[warp] = (u64)incl;
        s_hi[warp] = (u64)((u128)incl >> 64);
    }
    __syncthreads();
    i128 woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        i128 t = (i128)(((u128)s_hi[w] << 64) | s_lo[w]);
        if (w < (int)warp) woff += t;
        total += t;
    }
    i128 excl = woff + incl - v;
    if (PHASE == 0) {
        if (threadIdx.x == 0) blocksum[b] = total;
    } else {
        if (p <= n) wide[p] = base[b] + excl;
    }
}

// coordinate of the rights that carry a score, 0 for the others (start order: perm == nullptr)
__global__ void k_masked_coords(const u32 *__restrict__ coord, const u32 *__restrict__ vbits_a, const u32 *__restrict__ perm, size_t n,
                                u32 *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 j = perm ? perm[i] : (u32)i;
    out[i] = ((vbits_a[j >> 5] >> (j & 31)) & 1u) ? coord[i] : 0u;
}

// Exclusive scan of nb i128 block sums by one CTA (nb = n/256 + 1: small).
__global__ void __launch_bounds__(1024) k_scan_i128_single(const i128 *__restrict__ in0, size_t nb, i128 *__restrict__ out0,
                                                           const i128 *__restrict__ in1, i128 *__restrict__ out1) {
    // (CTA 0 scans in0 -> out0; a second CTA, if launched, in1 -> out1)
    const i128 *__restrict__ in = blockIdx.x ? in1 : in0;
    i128 *__restrict__ out = blockIdx.x ? out1 : out0;
    __shared__ u64 s_lo[32], s_hi[32];
    const size_t per = (nb + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per;
    const size_t hi = lo + per < nb ? lo + per : nb;
    i128 t = 0;
    for (size_t i = lo; i < hi; i++) t += in[i];
    // inclusive scan of the 1024 thread sums: warp shuffles, then the 32 warp totals by warp 0
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    i128 incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 tlo = __shfl_up_sync(GRB_FULL, (u64)incl, o);
        const u64 thi = __shfl_up_sync(GRB_FULL, (u64)((u128)incl >> 64), o);
        if (lane >= (u32)o) incl += (i128)(((u128)thi << 64) | tlo);
    }
    if (lane == 31) {
        s_lo[warp] = (u64)incl;
        s_hi[warp] = (u64)((u128)incl >> 64);
    }
    __syncthreads();
    if (warp == 0) {
        i128 w = (i128)(((u128)s_hi[lane] << 64) | s_lo[lane]);
        i128 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 tlo = __shfl_up_sync(GRB_FULL, (u64)wi, o);
            const u64 thi = __shfl_up_sync(GRB_FULL, (u64)((u128)wi >> 64), o);
            if (lane >= (u32)o) wi += (i128)(((u128)thi << 64) | tlo);
        }
        const i128 wex = wi - w;  // exclusive prefix of the warp totals
        s_lo[lane] = (u64)wex;
        s_hi[lane] = (u64)((u128)wex >> 64);
    }
    __syncthreads();
    i128 run = (i128)(((u128)s_hi[warp] << 64) | s_lo[warp]) + incl - t;
    for (size_t i = lo; i < hi; i++) {
        const i128 c = in[i];
        out[i] = run;
        run += c;
    }
}

// ---- member-sweep tables ---------------------------------------------------------

// (order-preserving score image, position) pairs; None sorts last
__global__ void k_rank_keys(const double *__restrict__ score_a, const u32 *__restrict__ vbits_a, size_t n, u64 *__restrict__ keys,
                            u32 *__restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = vbits_a ? ((vbits_a[i >> 5] >> (i & 31)) & 1u) : true;
    keys[i] = ok ? f64_key(score_a[i]) : ~0ull;
    vals[i] = (u32)i;
}

__global__ void k_rank_scatter(const u64 *__restrict__ keys, const u32 *__restrict__ vals, size_t n, u32 n_valid,
                               u32 *__restrict__ rank_a, double *__restrict__ score_sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + PAD) return;
    if (i >= n) {
        rank_a[i] = GRB_RANK_NONE;
        return;
    }
    if (i < n_valid) {
        rank_a[vals[i]] = (u32)i;
        score_sorted[i] = key_f64(keys[i]);
    } else {
        rank_a[vals[i]] = GRB_RANK_NONE;
    }
}

// inclusive prefix maximum of nb values by one CTA
__global__ void __launch_bounds__(1024) k_prefix_max_single(const u32 *__restrict__ in, size_t nb, u32 *__restrict__ out) {
    __shared__ u32 s_m[1024];
    const size_t per = (nb + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per;
    const size_t hi = lo + per < nb ? lo + per : nb;
    u32 m = 0;
    for (size_t i = lo; i < hi; i++) m = max(m, in[i]);
    s_m[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (int k = 0; k < 1024; k++) {
            u32 c = s_m[k];
            s_m[k] = run;
            run = max(run, c);
        }
    }
    __syncthreads();
    u32 run = s_m[threadIdx.x];
    for (size_t i = lo; i < hi; i++) {
        run = max(run, in[i]);
        out[i] = run;
    }
}

// fb[b] = first position j < 64 b with ends[j] > starts[64 b], else 64 b.  Every right that straddles a
// coordinate x >= starts[64 b] (start < x < end) sits at or after fb[b]: a straddler of x that starts before
// starts[64 b] ends after it, and one that starts later is positioned later.
__global__ void k_first_back(const u32 *__restrict__ starts, const u32 *__restrict__ ends, const u32 *__restrict__ pmax, size_t nb,
                             u32 *__restrict__ fb) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const u32 x = starts[b << GRB_BME_SHIFT];
    size_t lo = 0, hi = b;  // first block b' in [0, b) with pmax[b'] > x
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (pmax[mid] > x)
            hi = mid;
        else
            lo = mid + 1;
    }
    u32 r = (u32)(b << GRB_BME_SHIFT);
    if (lo < b) {
        const size_t j0 = lo << GRB_BME_SHIFT;
        for (int k = 0; k < 64; k++)
            if (ends[j0 + k] > x) {
                r = (u32)(j0 + k);
                break;
            }
    }
    fb[b] = r;
}

// ---- MIN / MAX without a sweep ------------------------------------------------------------
// A left's group is [pp, ub) in start order (range minimum: block sparse table) plus the rights that straddle
// l.start (start < x < end, x = l.start).  As x grows the straddler set changes at "open" events (coordinate
// start + 1) and "close" events (coordinate end); after all events with coordinate <= x there have been
// pp(x) = #{start < x} opens and lb(x) = #{end <= x} closes, so the set -- hence its smallest / largest rank --
// is a function of t = pp + lb alone.  Right r is in the set for t in [P_open + 1, P_close + 1) (P = 0-based
// position of its event in the merged order, closes first at equal coordinates).  stab_min / stab_max are
// filled by an offline range-chmin: each range drops its rank into two overlapping power-of-two cells of a
// dual sparse table, which is then pushed down level by level.

__global__ void k_block_rank_minmax(const u32 *__restrict__ rank_a, size_t n, u32 *__restrict__ bmin, int *__restrict__ bmax) {
    const size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nb = (n + 31) >> 5;
    if (b >= nb) return;
    const size_t i = (b << 5) + lane_id();
    const u32 r = i < n ? rank_a[i] : GRB_RANK_NONE;
    const u32 mn = __reduce_min_sync(GRB_FULL, r);
    const int mx = __reduce_max_sync(GRB_FULL, (int)r);
    if (lane_id() == 0) {
        bmin[b] = mn;
        bmax[b] = mx;
    }
}

__global__ void k_sparse_level(const u32 *__restrict__ pmin, const int *__restrict__ pmax, u32 nb, u32 half, u32 *__restrict__ omin,
                               int *__restrict__ omax) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const u32 j = i + half < nb ? i + half : i;
    omin[i] = min(pmin[i], pmin[j]);
    omax[i] = max(pmax[i], pmax[j]);
}

// per right (by end-order position j): its life [lo, hi) in event time, and the longest life
__global__ void k_stab_ranges(const u32 *__restrict__ starts, const u32 *__restrict__ ends_sorted, const u32 *__restrict__ perm_e,
                              u32 n, u32 *__restrict__ lo_out, u32 *__restrict__ hi_out, u32 *max_len) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    u32 len = 0;
    if (j < n) {
        const u32 i = perm_e[j];
        const u32 oc = starts[i] + 1, cc = ends_sorted[j];
        // P_open = i + #{end <= open coordinate};  P_close = j + #{start + 1 < close coordinate} = j + #{start < cc - 1}
        u32 a = 0, b = n;
        while (a < b) {
            const u32 m = a + ((b - a) >> 1);
            if (ends_sorted[m] <= oc)
                a = m + 1;
            else
                b = m;
        }
        const u32 p_open = i + a;
        a = 0, b = n;
        while (a < b) {
            const u32 m = a + ((b - a) >> 1);
            if (starts[m] + 1 < cc)
                a = m + 1;
            else
                b = m;
        }
        const u32 p_close = j + a;
        const u32 lo = p_open + 1, hi = p_close + 1;
        lo_out[j] = lo;
        hi_out[j] = hi > lo ? hi : lo;
        len = hi > lo ? hi - lo : 0;
    }
    len = warp_max(len);
    if (lane_id() == 0 && len) atomicMax(max_len, len);
}

__global__ void k_stab_drop(const u32 *__restrict__ rank_a, const u32 *__restrict__ perm_e, const u32 *__restrict__ lo_in,
                            const u32 *__restrict__ hi_in, u32 n, size_t T, u32 *__restrict__ tmin, int *__restrict__ tmax) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u32 r = rank_a[perm_e[j]];
    const u32 lo = lo_in[j], hi = hi_in[j];
    if (r == GRB_RANK_NONE || hi <= lo) return;
    const int k = 31 - __clz(hi - lo);
    const size_t base = (size_t)k * T;
    atomicMin(&tmin[base + lo], r);
    atomicMin(&tmin[base + hi - (1u << k)], r);
    atomicMax(&tmax[base + lo], (int)r);
    atomicMax(&tmax[base + hi - (1u << k)], (int)r);
}

// level k -> level k - 1: a cell [s, s + 2^k) is the union of [s, s + half) and [s + half, s + 2^k)
__global__ void k_stab_push(const u32 *__restrict__ umin, const int *__restrict__ umax, size_t T, u32 half, u32 *__restrict__ dmin,
                            int *__restrict__ dmax) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    u32 mn = min(dmin[t], umin[t]);
    int mx = max(dmax[t], umax[t]);
    if (t >= half) {
        mn = min(mn, umin[t - half]);
        mx = max(mx, umax[t - half]);
    }
    dmin[t] = mn;
    dmax[t] = mx;
}

inline unsigned blocks_for(size_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

int dev_alloc(grb_ctx *ctx, void **p, size_t bytes) {
    GRB_CUDA(ctx, grb_dev_malloc(ctx, p, bytes));
    return GRB_OK;
}

void free_scores(grb_index *ix) {
    grb_dev_free(ix->ctx, ix->score_a);
    grb_dev_free(ix->ctx, ix->vbits_a);
    grb_dev_free(ix->ctx, ix->vcnt_a);
    grb_dev_free(ix->ctx, ix->vcnt_b);
    grb_dev_free(ix->ctx, ix->pfx_a);
    grb_dev_free(ix->ctx, ix->pfx_b);
    grb_dev_free(ix->ctx, ix->pfh_a);
    grb_dev_free(ix->ctx, ix->pfh_b);
    grb_dev_free(ix->ctx, ix->nf_a);
    grb_dev_free(ix->ctx, ix->nf_b);
    grb_dev_free(ix->ctx, ix->vc16_a);
    grb_dev_free(ix->ctx, ix->vc16_b);
    ix->pfh_a = ix->pfh_b = nullptr;
    ix->nf_a = ix->nf_b = nullptr;
    ix->vc16_a = ix->vc16_b = nullptr;
    grb_dev_free(ix->ctx, ix->base_a);
    grb_dev_free(ix->ctx, ix->base_b);
    grb_dev_free(ix->ctx, ix->rank_a);
    grb_dev_free(ix->ctx, ix->score_sorted);
    grb_dev_free(ix->ctx, ix->stab_min);
    grb_dev_free(ix->ctx, ix->stab_max);
    grb_dev_free(ix->ctx, ix->sp_min);
    grb_dev_free(ix->ctx, ix->sp_max);
    ix->stab_min = ix->sp_min = nullptr;
    ix->stab_max = ix->sp_max = nullptr;
    ix->minmax_ready = false;
    grb_dev_free(ix->ctx, ix->wsum_a);
    grb_dev_free(ix->ctx, ix->wsum_b);
    grb_dev_free(ix->ctx, ix->sum_starts_v);
    grb_dev_free(ix->ctx, ix->sum_ends_v);
    ix->wsum_a = ix->wsum_b = nullptr;
    ix->sum_starts_v = ix->sum_ends_v = nullptr;
    ix->wm_ready = ix->wm_unavailable = false;
    ix->rank_a = nullptr;
    ix->score_sorted = nullptr;
    ix->members_ready = false;
    ix->has_nonfinite = false;
    ix->n_valid = 0;
    ix->score_a = nullptr;
    ix->vbits_a = ix->vcnt_a = ix->vcnt_b = nullptr;
    ix->pfx_a = ix->pfx_b = nullptr;
    ix->base_a = ix->base_b = nullptr;
    ix->has_scores = ix->has_nulls = ix->has_nan = false;
    ix->sum_mode = GRB_SUM_SWEEP;
}

int build_prefix_both(grb_ctx *ctx, grb_index *ix) {
    const size_t n = ix->n;
    const size_t nb = (n >> GRB_BASE_SHIFT) + 1;
    TempBuf bsum, bbase;
    GRB_TRY(bsum.alloc(ctx, 2 * nb * sizeof(i128)));
    FxSide a, b;
    memset(&a, 0, sizeof(a));
    memset(&b, 0, sizeof(b));
    b.perm = ix->perm_e;
    a.blocksum = bsum.as<i128>();
    b.blocksum = bsum.as<i128>() + nb;
    k_fx_blocks<0><<<(unsigned)((2 * nb + 7) / 8), 256, 0, ctx->stream>>>(ix->score_a, n, ix->e0, nb, a, b);
    GRB_LAUNCHED(ctx);
    if (ix->sum_mode == GRB_SUM_COMPACT || ix->sum_mode == GRB_SUM_WIDE96) {
        GRB_TRY(dev_alloc(ctx, (void **)&ix->base_a, nb * sizeof(i128)));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->base_b, nb * sizeof(i128)));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->pfx_a, (n + 1 + PAD) * sizeof(u64)));  // padded: bulk copies round up to 16 B
        GRB_TRY(dev_alloc(ctx, (void **)&ix->pfx_b, (n + 1 + PAD) * sizeof(u64)));
        k_scan_i128_single<<<2, 1024, 0, ctx->stream>>>(a.blocksum, nb, ix->base_a, b.blocksum, ix->base_b);
        GRB_LAUNCHED(ctx);
        a.base = ix->base_a, b.base = ix->base_b;
        a.pfx = ix->pfx_a, b.pfx = ix->pfx_b;
        if (ix->sum_mode == GRB_SUM_COMPACT) {
            k_fx_blocks<1><<<(unsigned)((2 * nb + 7) / 8), 256, 0, ctx->stream>>>(ix->score_a, n, ix->e0, nb, a, b);
        } else {
            GRB_TRY(dev_alloc(ctx, (void **)&ix->pfh_a, (n + 1 + PAD) * sizeof(u32)));
            GRB_TRY(dev_alloc(ctx, (void **)&ix->pfh_b, (n + 1 + PAD) * sizeof(u32)));
            a.pfh = ix->pfh_a, b.pfh = ix->pfh_b;
            k_fx_blocks<3><<<(unsigned)((2 * nb + 7) / 8), 256, 0, ctx->stream>>>(ix->score_a, n, ix->e0, nb, a, b);
        }
        GRB_LAUNCHED(ctx);
    } else {
        GRB_TRY(bbase.alloc(ctx, 2 * nb * sizeof(i128)));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->base_a, (n + 1) * sizeof(i128)));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->base_b, (n + 1) * sizeof(i128)));
        k_scan_i128_single<<<2, 1024, 0, ctx->stream>>>(a.blocksum, nb, bbase.as<i128>(), b.blocksum, bbase.as<i128>() + nb);
        GRB_LAUNCHED(ctx);
        a.base = bbase.as<i128>(), b.base = bbase.as<i128>() + nb;
        a.wide = ix->base_a, b.wide = ix->base_b;
        k_fx_blocks<2><<<(unsigned)((2 * nb + 7) / 8), 256, 0, ctx->stream>>>(ix->score_a, n, ix->e0, nb, a, b);
        GRB_LAUNCHED(ctx);
    }
    return GRB_OK;
}

// prefix counts of the +inf / -inf / NaN scores in one order (perm == nullptr: start order)
int build_nf(grb_ctx *ctx, grb_index *ix, const u32 *perm, uint4 **out) {
    const size_t n = ix->n;
    TempBuf fl, c0, c1, c2;
    GRB_TRY(fl.alloc(ctx, n * 4));
    GRB_TRY(c0.alloc(ctx, (n + 1) * 4));
    GRB_TRY(c1.alloc(ctx, (n + 1) * 4));
    GRB_TRY(c2.alloc(ctx, (n + 1) * 4));
    TempBuf *cs[3] = {&c0, &c1, &c2};
    for (int which = 0; which < 3; which++) {
        k_nf_flags<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->score_a, ix->has_nulls ? ix->vbits_a : nullptr, perm, n, which, fl.as<u32>());
        GRB_LAUNCHED(ctx);
        GRB_TRY(grb_scan_u32_to_u32(ctx, fl.as<u32>(), cs[which]->as<u32>(), n, true));
    }
    GRB_TRY(dev_alloc(ctx, (void **)out, (n + 1) * sizeof(uint4)));
    k_nf_pack<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(c0.as<u32>(), c1.as<u32>(), c2.as<u32>(), n + 1, *out);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

int ceil_log2(u64 x) {
    int l = 0;
    while ((1ull << l) < x) l++;
    return l;
}

}  // namespace

extern "C" {

int grb_index_build(grb_ctx *ctx, const uint32_t *starts, const uint32_t *ends, const uint64_t *data_idx, size_t n,
                    grb_index **out) {
    if (!ctx || !out) return GRB_ERR_INVALID_ARG;
    *out = nullptr;
    if (n && (!starts || !ends)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "index_build: NULL range buffer");
    if (n >= 0xfffffff0ull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "index_build: more than 2^32-16 ranges on one seqname");
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    grb_index *ix = (grb_index *)calloc(1, sizeof(grb_index));
    if (!ix) return grb_fail(ctx, GRB_ERR_OOM, "index_build: host allocation failed");
    ix->ctx = ctx;
    ix->n = (u32)n;
    ix->sum_mode = GRB_SUM_SWEEP;
    int rc = GRB_OK;
    do {
        if ((rc = dev_alloc(ctx, (void **)&ix->starts, (n + PAD) * 4))) break;
        if ((rc = dev_alloc(ctx, (void **)&ix->ends, (n + PAD) * 4))) break;
        if ((rc = dev_alloc(ctx, (void **)&ix->ends_sorted, (n + PAD) * 4))) break;
        if ((rc = dev_alloc(ctx, (void **)&ix->didx, n * 4))) break;
        if ((rc = dev_alloc(ctx, (void **)&ix->perm_e, (n + PAD) * 4))) break;  // padded: k_sweep3 bulk-copies 16-byte multiples
        if ((rc = dev_alloc(ctx, (void **)&ix->bme, (((n + 63) >> GRB_BME_SHIFT) + PAD) * 4))) break;
        InBuf d_starts, d_ends, d_idx;
        if ((rc = d_starts.stage(ctx, starts, n * 4))) break;
        if ((rc = d_ends.stage(ctx, ends, n * 4))) break;
        if ((rc = d_idx.stage(ctx, data_idx, data_idx ? n * 8 : 0))) break;
        TempBuf k0, k1, v0, v1, fl;
        if ((rc = k0.alloc(ctx, n * 8))) break;
        if ((rc = k1.alloc(ctx, n * 8))) break;
        if ((rc = v0.alloc(ctx, n * 4))) break;
        if ((rc = v1.alloc(ctx, n * 4))) break;
        if ((rc = fl.alloc(ctx, sizeof(BuildFlags)))) break;
        cudaMemsetAsync(fl.p, 0, sizeof(BuildFlags), ctx->stream);
        u64 *ks = k0.as<u64>();
        u32 *vs = v0.as<u32>();
        const u32 *order = nullptr;
        TempBuf order_buf;
        if (n && data_idx) {
            // ties on (start, end) must come out in data-index order: pre-order the input by data index
            k_idx_keys<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_idx.as<u64>(), n, k0.as<u64>(), v0.as<u32>(), fl.as<BuildFlags>());
            GRB_LAUNCHED_BREAK(ctx, rc)
            bool sorted0;
            if ((rc = grb_sort_pairs(ctx, k0.as<u64>(), v0.as<u32>(), k1.as<u64>(), v1.as<u32>(), n, &ks, &vs, &sorted0))) break;
            if (!sorted0) {
                if ((rc = order_buf.alloc(ctx, n * 4))) break;
                cudaMemcpyAsync(order_buf.p, vs, n * 4, cudaMemcpyDeviceToDevice, ctx->stream);
                order = order_buf.as<u32>();
            }
        }
        if (n) {
            const unsigned mk_blocks = blocks_for(n, 256) < (unsigned)ctx->sm_count * 8u ? blocks_for(n, 256) : (unsigned)ctx->sm_count * 8u;
            k_make_keys<<<mk_blocks, 256, 0, ctx->stream>>>(d_starts.as<u32>(), d_ends.as<u32>(), order, n, k0.as<u64>(), v0.as<u32>(),
                                                           fl.as<BuildFlags>());
            GRB_LAUNCHED_BREAK(ctx, rc)
        }
        if ((rc = grb_sort_pairs(ctx, k0.as<u64>(), v0.as<u32>(), k1.as<u64>(), v1.as<u32>(), n, &ks, &vs, nullptr))) break;
        BuildFlags hf;
        if ((rc = grb_read_small(ctx, &hf, fl.p, sizeof(hf)))) break;
        if (hf.invalid_range) {
            rc = grb_fail(ctx, GRB_ERR_INVALID_RANGE,
                          "index_build: a range has end <= start or end > 2^31-1 (the reference panics: "
                          "ranges/mod.rs:109, ranges/coitrees.rs:53-54)");
            break;
        }
        if (hf.idx_too_big) {
            rc = grb_fail(ctx, GRB_ERR_UNSUPPORTED, "index_build: data index >= 2^32-1");
            break;
        }
        k_unpack_start_order<<<blocks_for(n + PAD, 256), 256, 0, ctx->stream>>>(ks, vs, d_idx.as<u64>(), n, ix->starts, ix->ends,
                                                                               ix->didx);
        GRB_LAUNCHED_BREAK(ctx, rc)
        {
            // coordinate bins of ~2 rights each (so one round of compares usually ends a search); the largest end bounds every key
            const u32 maxend = hf.max_end;
            int shift = 0;
            const char *ld = getenv("GRB_LUT_DIV");  // tuning knob: average rights per bin (default 2)
            const u64 lut_div = ld && atoi(ld) > 0 ? (u64)atoi(ld) : 2;
            while (shift < 31 && ((u64)maxend >> shift) > (u64)n / lut_div + 1) shift++;
            ix->lut_shift = shift;
            ix->lut_bmax = (maxend >> shift) + 1;
            const u32 bmax = ix->lut_bmax;
            size_t entries = (size_t)bmax + 1 + PAD;
            if ((rc = dev_alloc(ctx, (void **)&ix->lut_a, entries * 4))) break;
            if ((rc = dev_alloc(ctx, (void **)&ix->lut_b, entries * 4))) break;
            if ((rc = dev_alloc(ctx, (void **)&ix->lut16_a, entries * 2))) break;
            if ((rc = dev_alloc(ctx, (void **)&ix->lut16_b, entries * 2))) break;
            k_build_lut<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(ix->starts, (u32)n, shift, bmax, ix->lut_a, ix->lut16_a);
            GRB_LAUNCHED_BREAK(ctx, rc)
            // end order + the ends' table in one go: counting sort over the bins (see k_bin_count)
            TempBuf cnt;
            if ((rc = cnt.alloc(ctx, ((size_t)bmax + 1) * 4))) break;
            cudaMemsetAsync(cnt.p, 0, ((size_t)bmax + 1) * 4, ctx->stream);
            if (n) {
                k_bin_count<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->ends, (u32)n, shift, bmax, cnt.as<u32>());
                GRB_LAUNCHED_BREAK(ctx, rc)
            }
            if ((rc = grb_scan_u32_to_u32(ctx, cnt.as<u32>(), ix->lut_b, (size_t)bmax + 1, true))) break;  // lut_b[bmax + 1] = n
            k_fill_u32<<<1, PAD, 0, ctx->stream>>>(ix->lut_b + bmax + 2, PAD - 1, (u32)n);
            GRB_LAUNCHED_BREAK(ctx, rc)
            k_narrow_u16<<<blocks_for(entries, 256), 256, 0, ctx->stream>>>(ix->lut_b, entries, ix->lut16_b);
            GRB_LAUNCHED_BREAK(ctx, rc)
            if (n) {
                cudaMemsetAsync(cnt.p, 0, ((size_t)bmax + 1) * 4, ctx->stream);
                // (the sort buffers are free again: the one that does not hold the sorted keys takes the bins in arrival order)
                u32 *tmp_e = (u32 *)((ks == k0.as<u64>()) ? k1.as<u64>() : k0.as<u64>());
                u32 *tmp_p = (vs == v0.as<u32>()) ? v1.as<u32>() : v0.as<u32>();
                k_bin_scatter<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->ends, (u32)n, shift, bmax, ix->lut_b, cnt.as<u32>(), tmp_e, tmp_p);
                GRB_LAUNCHED_BREAK(ctx, rc)
                k_bin_place<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->lut_b, shift, bmax, tmp_e, tmp_p, (u32)n, ix->ends_sorted, ix->perm_e,
                                                                        fl.as<BuildFlags>());
                GRB_LAUNCHED_BREAK(ctx, rc)
            }
            k_pad_keys<<<1, PAD, 0, ctx->stream>>>(ix->ends_sorted + n);
            GRB_LAUNCHED_BREAK(ctx, rc)
        }
        if (n) {
            size_t nbw = (n + 63) >> GRB_BME_SHIFT;
            k_block_max_end<<<blocks_for(nbw * 32, 256), 256, 0, ctx->stream>>>(ix->ends, n, ix->bme);
            GRB_LAUNCHED_BREAK(ctx, rc)
        }
        if ((rc = grb_read_small(ctx, &hf, fl.p, sizeof(hf)))) break;
        if (hf.heavy_bin) {
            // a coordinate bin with hundreds of ends (e.g. thousands of rights ending on one base): the stable radix sort
            // gives the same (end, position) order
            u64 *k2 = (ks == k0.as<u64>()) ? k1.as<u64>() : k0.as<u64>();
            u32 *v2 = (vs == v0.as<u32>()) ? v1.as<u32>() : v0.as<u32>();
            if ((rc = grb_argsort_u32(ctx, ix->ends, ix->ends_sorted, ix->perm_e, (u32 *)k2, v2, n))) break;
            const cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                rc = grb_fail(ctx, GRB_ERR_CUDA, "index_build: %s", cudaGetErrorString(e));
                break;
            }
        }
    } while (0);
    if (rc != GRB_OK) {
        grb_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return GRB_OK;
}

void grb_index_destroy(grb_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    free_scores(ix);
    grb_dev_free(ix->ctx, ix->starts);
    grb_dev_free(ix->ctx, ix->ends);
    grb_dev_free(ix->ctx, ix->didx);
    grb_dev_free(ix->ctx, ix->bme);
    grb_dev_free(ix->ctx, ix->ends_sorted);
    grb_dev_free(ix->ctx, ix->perm_e);
    grb_dev_free(ix->ctx, ix->lut_a);
    grb_dev_free(ix->ctx, ix->lut_b);
    grb_dev_free(ix->ctx, ix->lut16_a);
    grb_dev_free(ix->ctx, ix->lut16_b);
    grb_dev_free(ix->ctx, ix->fb);
    grb_dev_free(ix->ctx, ix->sum_starts);
    grb_dev_free(ix->ctx, ix->sum_ends);
    free(ix);
}

size_t grb_index_len(const grb_index *ix) { return ix ? ix->n : 0; }

int grb_index_exact_sums(const grb_index *ix) { return ix && ix->has_scores && ix->sum_mode != GRB_SUM_SWEEP; }

int grb_index_set_scores(grb_index *ix, const double *by_data_idx, const uint8_t *valid, size_t n_data) {
    if (!ix) return GRB_ERR_INVALID_ARG;
    grb_ctx *ctx = ix->ctx;
    if (n_data && !by_data_idx) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "set_scores: NULL scores");
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    free_scores(ix);
    const size_t n = ix->n;
    ix->n_data = n_data;
    InBuf d_scores, d_valid;
    GRB_TRY(d_scores.stage(ctx, by_data_idx, n_data * 8));
    GRB_TRY(d_valid.stage(ctx, valid, valid ? n_data : 0));
    GRB_TRY(dev_alloc(ctx, (void **)&ix->score_a, n * 8));
    TempBuf stb, vflag_a, vflag_b;
    GRB_TRY(stb.alloc(ctx, sizeof(ScoreStats)));
    ScoreStats init;
    memset(&init, 0, sizeof(init));
    init.min_lsb_exp = 1 << 30;
    init.max_exp = -(1 << 30);
    GRB_CUDA(ctx, cudaMemcpyAsync(stb.p, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (valid) {
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vbits_a, ((n + 31) / 32) * 4));
        GRB_TRY(vflag_a.alloc(ctx, n * 4));
    }
    if (n) {
        k_gather_scores<<<blocks_for(n, 1024), 256, 0, ctx->stream>>>(ix->didx, n, d_scores.as<double>(), d_valid.as<u8>(), n_data,
                                                                    ix->score_a, ix->vbits_a, valid ? vflag_a.as<u32>() : nullptr,
                                                                    stb.as<ScoreStats>());
        GRB_LAUNCHED(ctx);
    }
    ScoreStats st;
    GRB_TRY(grb_read_small(ctx, &st, stb.p, sizeof(st)));
    if (st.idx_out_of_range) {
        free_scores(ix);
        return grb_fail(ctx, GRB_ERR_INDEX_RANGE, "set_scores: a right range's data index is >= n_data (%zu)", n_data);
    }
    ix->has_scores = true;
    ix->has_nulls = st.n_null != 0;
    ix->has_nan = st.has_nan != 0;
    ix->has_nonfinite = st.has_nonfinite != 0;
    ix->n_valid = (u32)(n - st.n_null);
    if (ix->has_nulls) {
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vcnt_a, (n + 1) * 4));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vcnt_b, (n + 1) * 4));
        GRB_TRY(vflag_b.alloc(ctx, n * 4));
        k_permute_u32<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(vflag_a.as<u32>(), ix->perm_e, n, vflag_b.as<u32>());
        GRB_LAUNCHED(ctx);
        GRB_TRY(grb_scan_u32_to_u32(ctx, vflag_a.as<u32>(), ix->vcnt_a, n, true));
        GRB_TRY(grb_scan_u32_to_u32(ctx, vflag_b.as<u32>(), ix->vcnt_b, n, true));
    } else if (ix->has_nonfinite) {
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vcnt_a, (n + 1) * 4));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vcnt_b, (n + 1) * 4));
        k_iota_u32<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(n + 1, ix->vcnt_a, ix->vcnt_b);
        GRB_LAUNCHED(ctx);
    }
    if (ix->has_nulls || ix->has_nonfinite) {
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vc16_a, (n + 1 + PAD) * 2));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->vc16_b, (n + 1 + PAD) * 2));
        k_narrow_u16<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(ix->vcnt_a, n + 1, ix->vc16_a);
        GRB_LAUNCHED(ctx);
        k_narrow_u16<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(ix->vcnt_b, n + 1, ix->vc16_b);
        GRB_LAUNCHED(ctx);
    }
    if (!ix->has_nulls && ix->vbits_a) {
        grb_dev_free(ix->ctx, ix->vbits_a);
        ix->vbits_a = nullptr;
    }
    // choose how SUM / MEAN are answered
    ix->sum_mode = GRB_SUM_SWEEP;
    ix->e0 = 0;
    const char *force = getenv("GRB_SUM_MODE");  // debug: 0 sweep, 2 wide, 3 wide96 (only if representable)
    // (non-finite scores count as 0 in the fixed point; the nf_* counters below carry them)
    if (!st.has_subnormal) {
        int W = 1, e0 = 0, emax = 0;
        if (st.n_nonzero) {
            e0 = st.min_lsb_exp;
            emax = st.max_exp;
            W = emax - e0 + 1;
        }
        int L = ceil_log2((u64)n + 1);
        bool representable = (W + L <= 126) && (emax + L + 1 < 1023) && (e0 >= -1022);
        if (representable) {
            ix->e0 = e0;
            ix->width = W;
            ix->sum_mode = (W + GRB_BASE_SHIFT <= 62) ? GRB_SUM_COMPACT : (W + GRB_BASE_SHIFT <= 94) ? GRB_SUM_WIDE96 : GRB_SUM_WIDE;
            if (force) {
                int f = atoi(force);
                if (f == GRB_SUM_SWEEP) ix->sum_mode = GRB_SUM_SWEEP;
                if (f == GRB_SUM_WIDE) ix->sum_mode = GRB_SUM_WIDE;
                if (f == GRB_SUM_WIDE96 && W + GRB_BASE_SHIFT <= 94) ix->sum_mode = GRB_SUM_WIDE96;
            }
        }
    }
    if (ix->sum_mode != GRB_SUM_SWEEP) {
        GRB_TRY(build_prefix_both(ctx, ix));
        if (ix->has_nonfinite && n) {
            GRB_TRY(build_nf(ctx, ix, nullptr, &ix->nf_a));
            GRB_TRY(build_nf(ctx, ix, ix->perm_e, &ix->nf_b));
        }
    }
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

// Build the indexes of a whole right-hand GRanges (one per seqname).  A build is a chain of small kernels with a
// few host round trips (sortedness, digit masks, score statistics), so one seqname cannot fill the GPU or the
// PCIe link on its own: up to four host threads, each with its own stream (a private clone of the context),
// take the seqnames round-robin, and the uploads and kernels of different seqnames overlap.  The finished
// indexes are handed to the caller's context.
int grb_index_build_batch(grb_ctx *ctx, int nseq, const uint32_t *const *starts, const uint32_t *const *ends,
                          const size_t *n, const double *const *scores, const uint8_t *const *valid, grb_index **out) {
    if (!ctx || nseq < 0 || (nseq && (!starts || !ends || !n || !out))) return GRB_ERR_INVALID_ARG;
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int s = 0; s < nseq; s++) out[s] = nullptr;
    if (nseq == 0) return GRB_OK;
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // earlier work on the caller's stream precedes the workers'
    int wmax = 4;  // GRB_BUILD_WORKERS: host threads (private streams) of a batched build
    if (const char *bw = getenv("GRB_BUILD_WORKERS")) wmax = atoi(bw) > 0 && atoi(bw) <= 15 ? atoi(bw) : 4;
    const int nworkers = nseq < wmax ? nseq : wmax;
    std::vector<grb_ctx> clones((size_t)nworkers);
    std::vector<int> rcs((size_t)nworkers, GRB_OK);
    for (int w = 0; w < nworkers; w++) {
        clones[w] = *ctx;  // shares the device, the pool and the zero / flag words
        clones[w].ctl_slot = w + 1;
        clones[w].batch_dev = nullptr;
        clones[w].batch_host = nullptr;
        clones[w].batch_bytes = clones[w].batch_cap = 0;
        clones[w].launches = 0;
        clones[w].err[0] = 0;
        if (cudaStreamCreateWithFlags(&clones[w].own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            for (int v = 0; v < w; v++) cudaStreamDestroy(clones[v].own_stream);
            return grb_fail(ctx, GRB_ERR_CUDA, "index_build_batch: cannot create a stream");
        }
        clones[w].stream = clones[w].own_stream;
    }
    auto work = [&](int w) {
        grb_ctx *c = &clones[w];
        cudaSetDevice(c->device);
        for (int s = w; s < nseq; s += nworkers) {
            int rc = grb_index_build(c, starts[s], ends[s], nullptr, n[s], &out[s]);
            if (rc == GRB_OK && scores && scores[s]) rc = grb_index_set_scores(out[s], scores[s], (valid && valid[s]) ? valid[s] : nullptr, n[s]);
            if (rc != GRB_OK) {
                rcs[w] = rc;
                return;
            }
        }
    };
    {
        std::vector<std::thread> threads;
        for (int w = 1; w < nworkers; w++) threads.emplace_back(work, w);
        work(0);
        for (auto &t : threads) t.join();
    }
    int rc = GRB_OK;
    for (int w = 0; w < nworkers; w++) {
        cudaStreamSynchronize(clones[w].own_stream);
        cudaStreamDestroy(clones[w].own_stream);
        ctx->launches += clones[w].launches;
        if (rcs[w] != GRB_OK && rc == GRB_OK) {
            rc = rcs[w];
            memcpy(ctx->err, clones[w].err, sizeof(ctx->err));
        }
    }
    for (int s = 0; s < nseq; s++)
        if (out[s]) out[s]->ctx = ctx;  // the clones go away; the indexes belong to the caller's context
    if (rc != GRB_OK) {
        for (int s = 0; s < nseq; s++) {
            grb_index_destroy(out[s]);
            out[s] = nullptr;
        }
    }
    return rc;
}

}  // extern "C"

// Tables of the staged member sweep (k_sweep3), built the first time MIN / MAX / MEDIAN are asked of an index:
// the rank of every score (one more radix sort), the scores in rank order, and the reach-back table.
int grb_index_prepare_members(grb_ctx *ctx, grb_index *ix) {
    if (!ix || ix->members_ready || !ix->n || !ix->has_scores) return GRB_OK;
    const size_t n = ix->n;
    GRB_TRY(dev_alloc(ctx, (void **)&ix->rank_a, (n + PAD) * 4));
    GRB_TRY(dev_alloc(ctx, (void **)&ix->score_sorted, ((size_t)ix->n_valid + 1) * 8));
    {
        TempBuf k0, k1, v0, v1;
        GRB_TRY(k0.alloc(ctx, n * 8));
        GRB_TRY(k1.alloc(ctx, n * 8));
        GRB_TRY(v0.alloc(ctx, n * 4));
        GRB_TRY(v1.alloc(ctx, n * 4));
        k_rank_keys<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->score_a, ix->has_nulls ? ix->vbits_a : nullptr, n, k0.as<u64>(),
                                                                v0.as<u32>());
        GRB_LAUNCHED(ctx);
        u64 *ks;
        u32 *vs;
        GRB_TRY(grb_sort_pairs(ctx, k0.as<u64>(), v0.as<u32>(), k1.as<u64>(), v1.as<u32>(), n, &ks, &vs, nullptr));
        k_rank_scatter<<<blocks_for(n + PAD, 256), 256, 0, ctx->stream>>>(ks, vs, n, ix->n_valid, ix->rank_a, ix->score_sorted);
        GRB_LAUNCHED(ctx);
    }
    if (!ix->fb) {
        const size_t nb = (n + 63) >> GRB_BME_SHIFT;
        TempBuf pm;
        GRB_TRY(pm.alloc(ctx, nb * 4));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->fb, nb * 4));
        k_prefix_max_single<<<1, 1024, 0, ctx->stream>>>(ix->bme, nb, pm.as<u32>());
        GRB_LAUNCHED(ctx);
        k_first_back<<<blocks_for(nb, 128), 128, 0, ctx->stream>>>(ix->starts, ix->ends, pm.as<u32>(), nb, ix->fb);
        GRB_LAUNCHED(ctx);
    }
    ix->members_ready = true;
    return GRB_OK;
}

int grb_index_prepare_minmax(grb_ctx *ctx, grb_index *ix) {
    if (!ix || ix->minmax_ready || !ix->n || !ix->has_scores) return GRB_OK;
    GRB_TRY(grb_index_prepare_members(ctx, ix));
    const size_t n = ix->n;
    // ---- range minimum / maximum over rank_a: per-32-block values and their sparse table
    const u32 nb = (u32)((n + 31) >> 5);
    int levels = 1;
    while ((1u << levels) <= nb) levels++;
    ix->sp_nb = nb;
    ix->sp_levels = levels;
    GRB_TRY(dev_alloc(ctx, (void **)&ix->sp_min, (size_t)levels * nb * 4));
    GRB_TRY(dev_alloc(ctx, (void **)&ix->sp_max, (size_t)levels * nb * 4));
    k_block_rank_minmax<<<blocks_for((size_t)nb * 32, 256), 256, 0, ctx->stream>>>(ix->rank_a, n, ix->sp_min, ix->sp_max);
    GRB_LAUNCHED(ctx);
    for (int k = 1; k < levels; k++) {
        k_sparse_level<<<blocks_for(nb, 256), 256, 0, ctx->stream>>>(ix->sp_min + (size_t)(k - 1) * nb, ix->sp_max + (size_t)(k - 1) * nb, nb,
                                                                    1u << (k - 1), ix->sp_min + (size_t)k * nb, ix->sp_max + (size_t)k * nb);
        GRB_LAUNCHED(ctx);
    }
    // ---- stabbing minimum / maximum over event time t in [0, 2n]
    const size_t T = 2 * n + 1;
    TempBuf lo, hi, mx;
    GRB_TRY(lo.alloc(ctx, n * 4));
    GRB_TRY(hi.alloc(ctx, n * 4));
    GRB_TRY(mx.alloc(ctx, 4));
    GRB_CUDA(ctx, cudaMemsetAsync(mx.p, 0, 4, ctx->stream));
    k_stab_ranges<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->starts, ix->ends_sorted, ix->perm_e, (u32)n, lo.as<u32>(), hi.as<u32>(),
                                                              mx.as<u32>());
    GRB_LAUNCHED(ctx);
    u32 max_len = 0;
    GRB_CUDA(ctx, cudaMemcpyAsync(&max_len, mx.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int K = 0;
    while ((2u << K) <= max_len && K < 31) K++;  // floor(log2(max_len)), 0 when no right straddles anything
    GRB_TRY(dev_alloc(ctx, (void **)&ix->stab_min, T * 4));
    GRB_TRY(dev_alloc(ctx, (void **)&ix->stab_max, T * 4));
    {
        // levels 1..K are scratch; level 0 is the table itself
        TempBuf tmin, tmax;
        GRB_TRY(tmin.alloc(ctx, (size_t)(K + 1) * T * 4));
        GRB_TRY(tmax.alloc(ctx, (size_t)(K + 1) * T * 4));
        GRB_CUDA(ctx, cudaMemsetAsync(tmin.p, 0xff, (size_t)(K + 1) * T * 4, ctx->stream));  // GRB_RANK_NONE
        GRB_CUDA(ctx, cudaMemsetAsync(tmax.p, 0xff, (size_t)(K + 1) * T * 4, ctx->stream));  // -1
        k_stab_drop<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->rank_a, ix->perm_e, lo.as<u32>(), hi.as<u32>(), (u32)n, T, tmin.as<u32>(),
                                                                tmax.as<int>());
        GRB_LAUNCHED(ctx);
        for (int k = K; k >= 1; k--) {
            k_stab_push<<<blocks_for(T, 256), 256, 0, ctx->stream>>>(tmin.as<u32>() + (size_t)k * T, tmax.as<int>() + (size_t)k * T, T, 1u << (k - 1),
                                                                    tmin.as<u32>() + (size_t)(k - 1) * T, tmax.as<int>() + (size_t)(k - 1) * T);
            GRB_LAUNCHED(ctx);
        }
        GRB_CUDA(ctx, cudaMemcpyAsync(ix->stab_min, tmin.p, T * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        GRB_CUDA(ctx, cudaMemcpyAsync(ix->stab_max, tmax.p, T * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ix->minmax_ready = true;
    return GRB_OK;
}

// sum_starts[i] / sum_ends[i] = sum of the i smallest starts / ends (exact: below 2^63).  With them the overlap
// bp of a left with ALL rights is F(l.end) - F(l.start), F(x) = sum_r |[r.start, r.end) & [0, x)|
//   = sum_ends[#{end <= x}] - sum_starts[#{start < x}] + x * (#{start < x} - #{end <= x}):  no member is visited.
int grb_index_prepare_bp(grb_ctx *ctx, grb_index *ix) {
    if (!ix || ix->sum_starts || !ix->n) return GRB_OK;
    const size_t n = ix->n;
    GRB_TRY(dev_alloc(ctx, (void **)&ix->sum_starts, (n + 1) * 8));
    GRB_TRY(dev_alloc(ctx, (void **)&ix->sum_ends, (n + 1) * 8));
    GRB_TRY(grb_scan_u32_to_u64(ctx, ix->starts, ix->sum_starts, n, true));
    GRB_TRY(grb_scan_u32_to_u64(ctx, ix->ends_sorted, ix->sum_ends, n, true));
    return GRB_OK;
}

// WEIGHTED_MEAN in closed form needs exact prefixes of score * start (start order) and score * end (end order):
// sum_r score_r |l & r| = G(l.end) - G(l.start),  G(x) = WE[#{end <= x}] - WS[#{start < x}] + x (PA[#{start < x}] - PB[#{end <= x}]).
static int build_wprefix(grb_ctx *ctx, grb_index *ix, const u32 *perm, const u32 *coord, i128 **out) {
    const size_t n = ix->n;
    const size_t nb = (n >> GRB_BASE_SHIFT) + 1;
    TempBuf bsum, bbase;
    GRB_TRY(bsum.alloc(ctx, nb * sizeof(i128)));
    GRB_TRY(bbase.alloc(ctx, nb * sizeof(i128)));
    GRB_TRY(dev_alloc(ctx, (void **)out, (n + 1) * sizeof(i128)));
    k_wfx_blocks<0><<<(unsigned)nb, 256, 0, ctx->stream>>>(ix->score_a, perm, coord, n, ix->e0, bsum.as<i128>(), nullptr, nullptr);
    GRB_LAUNCHED(ctx);
    k_scan_i128_single<<<1, 1024, 0, ctx->stream>>>(bsum.as<i128>(), nb, bbase.as<i128>(), nullptr, nullptr);
    GRB_LAUNCHED(ctx);
    k_wfx_blocks<2><<<(unsigned)nb, 256, 0, ctx->stream>>>(ix->score_a, perm, coord, n, ix->e0, nullptr, bbase.as<i128>(), *out);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

int grb_index_prepare_wm(grb_ctx *ctx, grb_index *ix) {
    if (!ix || ix->wm_ready || ix->wm_unavailable || !ix->n || !ix->has_scores) return GRB_OK;
    const int L = ceil_log2((u64)ix->n + 1);
    // score (W bits) * coordinate (31 bits) summed over n rights must fit the signed 128-bit fixed point, and the
    // rounded result a double
    if (ix->sum_mode == GRB_SUM_SWEEP || ix->has_nonfinite || ix->width + L + 32 > 126 || ix->e0 + ix->width + L + 33 >= 1023) {
        ix->wm_unavailable = true;
        return GRB_OK;
    }
    const size_t n = ix->n;
    GRB_TRY(grb_index_prepare_bp(ctx, ix));
    GRB_TRY(build_wprefix(ctx, ix, nullptr, ix->starts, &ix->wsum_a));
    GRB_TRY(build_wprefix(ctx, ix, ix->perm_e, ix->ends_sorted, &ix->wsum_b));
    if (ix->has_nulls) {
        TempBuf m;
        GRB_TRY(m.alloc(ctx, n * 4));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->sum_starts_v, (n + 1) * 8));
        GRB_TRY(dev_alloc(ctx, (void **)&ix->sum_ends_v, (n + 1) * 8));
        k_masked_coords<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->starts, ix->vbits_a, nullptr, n, m.as<u32>());
        GRB_LAUNCHED(ctx);
        GRB_TRY(grb_scan_u32_to_u64(ctx, m.as<u32>(), ix->sum_starts_v, n, true));
        k_masked_coords<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ix->ends_sorted, ix->vbits_a, ix->perm_e, n, m.as<u32>());
        GRB_LAUNCHED(ctx);
        GRB_TRY(grb_scan_u32_to_u64(ctx, m.as<u32>(), ix->sum_ends_v, n, true));
    }
    ix->wm_ready = true;
    return GRB_OK;
}

extern "C" {

int grb_index_copy_sorted(const grb_index *ix, uint32_t *starts, uint32_t *ends, uint64_t *data_idx) {
    if (!ix) return GRB_ERR_INVALID_ARG;
    grb_ctx *ctx = ix->ctx;
    const size_t n = ix->n;
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (starts) GRB_CUDA(ctx, cudaMemcpyAsync(starts, ix->starts, n * 4, cudaMemcpyDefault, ctx->stream));
    if (ends) GRB_CUDA(ctx, cudaMemcpyAsync(ends, ix->ends, n * 4, cudaMemcpyDefault, ctx->stream));
    if (data_idx) {
        u32 *tmp = (u32 *)malloc(n * 4 + 4);
        if (!tmp) return grb_fail(ctx, GRB_ERR_OOM, "copy_sorted: host allocation failed");
        cudaError_t e = cudaMemcpyAsync(tmp, ix->didx, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            free(tmp);
            return grb_fail(ctx, GRB_ERR_CUDA, "copy_sorted: %s", cudaGetErrorString(e));
        }
        for (size_t i = 0; i < n; i++) data_idx[i] = tmp[i];
        free(tmp);
    }
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

}  // extern "C"
