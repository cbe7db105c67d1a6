// tilelib.cuh -- what the tile kernels (query.cu) and the pipelined map kernel (pipe.cu) share: tile geometry,
// the mbarrier / TMA bulk-copy PTX, the in-window searches, and the host-side batch descriptor.
#pragma once
#include <vector>

#include "common.cuh"

struct TileOut {
    u32 *counts;
    u8 *keep;
    unsigned long long *n_kept;
    int anti;
    u32 *ub, *lb, *pp;  // M_UBLB: #{start < l.end}, #{end <= l.start}, #{start < l.start}
    double *out;
    u8 *out_valid;
    int k;
    int need_sum;
    int ops[GRB_MAX_OPS];
    u32 *flags;  // context flag word (bit 0: invalid left range seen)
    u32 *stats;  // pipelined kernel: stats[16 + blockIdx.x] = tiles this CTA could not stage (mapped host words; may be NULL)
    u32 stats_gen;  // ... and stats[1] = this number, by CTA 0, after its count
    uint4 *planar;  // PIPE_UBLB beside k_sweep3: {sum lo, sum hi, count, scored members} per left here instead of the prefix-sum
                    // columns in `out` -- the sweep writes whole rows, nobody read-modify-writes partial sectors (may be NULL)
};

// ---- host side: one batch = the seqnames of one call, concatenated in GenomeMap order -------------------

struct Batch {
    const SeqDev *seqs = nullptr;
    const u32 *tile_begin = nullptr;
    u32 ntiles = 0;
    u64 total = 0;
    int nseq = 0;
    bool uniform = true;   // every present seqname: scores, no None, finite, compact prefixes (the FAST tile kernels)
    bool pipeable = true;  // every present seqname: scores with exact prefix sums in planes the pipeline can stage
    std::vector<u32> host_tile_begin;  // nseq + 1
    std::vector<int> var;              // per seqname: PIPE_VAR_* bits the pipelined kernel needs for it (-1: absent / empty)
};

enum { PIPE_VAR_NULLS = 1, PIPE_VAR_WIDE = 2 };

constexpr int PIPE_MULTI = 100;  // FAST value: several prefix-sum reductions (sum / sum-not-empty / mean / count) per left
constexpr int PIPE_UBLB = 102;   // FAST value: PIPE_MULTI for the columns the prefix sums answer (others are left alone) + the three
                                 // counts ub = #{start < l.end}, pp = #{start < l.start}, lb = #{end <= l.start} written out for the
                                 // member sweep / the closed forms that follow


namespace {

constexpr int TILE_TPB = 256;
constexpr int TILE_LPT = 4;
constexpr int TILE = TILE_TPB * TILE_LPT;
constexpr int TILE_SHIFT = 10;
static_assert(TILE == 1 << TILE_SHIFT, "tile size");
constexpr u32 WIN_CAP = 2048;  // keys per staged window (2 x 8 KB of shared memory)
constexpr u32 LUT_CAP = 1024;  // bin-table entries per staged slice (2 x 4 KB)

enum { M_COUNTS = 1, M_KEEP = 2, M_UBLB = 4, M_MAP = 8 };


// ---- PTX: mbarrier + TMA 1-D bulk copy ---------------------------------------

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// Same copy with an L2 evict-first policy: for data that is read once per pass (it should not push the
// prefix bases / descriptors / other CTAs' overlapping slices out of L2).
__device__ __forceinline__ u64 l2_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst, const void *src, u32 bytes, u64 *bar, u64 policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Poll, and sleep between polls: a waiting warp should not eat the issue slots the working warps need.
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    if (mbar_try_wait(bar, parity)) return;
    while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// ---- searches -------------------------------------------------------------------

// #keys < x in the answer interval [lo, hi] of a sorted array (shared or global).
__device__ __forceinline__ u32 bounded_lower_bound(const u32 *__restrict__ keys, u32 lo, u32 hi, u32 x) {
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (keys[mid] < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}


// #keys < x, scanning forward from a bin's first position `lo` (<= answer) four keys at a time.
// No upper bound is needed: `keys` is sorted and every key at or after the end of the searched
// slice is >= x (real keys of later bins, then the 0xffffffff sentinels that pad every key
// array), so the predicates are a run of trues followed by falses and the scan always stops.
// Bins hold ~4-8 keys; the four loads of a round are independent.
__device__ __forceinline__ u32 scan_lower_bound(const u32 *__restrict__ keys, u32 lo, u32 x) {
    while (true) {
        const u32 c = (keys[lo] < x ? 1u : 0u) + (keys[lo + 1] < x ? 1u : 0u) + (keys[lo + 2] < x ? 1u : 0u) +
                      (keys[lo + 3] < x ? 1u : 0u);
        if (c < 4) return lo + c;
        lo += 4;
    }
}

__device__ __forceinline__ int find_seq_by_tile(const u32 *__restrict__ tile_begin, int nseq, u32 t) {
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_begin[mid] <= t)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int find_seq_by_left(const SeqDev *__restrict__ seqs, int nseq, u64 g) {
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (seqs[mid].left_begin <= g)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Exact sum of the group [lb, ub) from the compact prefixes.  Small groups (the usual case): the
// sum fits 63 bits, so the difference of the prefixes mod 2^64 IS the sum; otherwise rebuild the
// two exact 128-bit prefixes from the per-256 bases.  Either way one rounding to f64.
__device__ __forceinline__ double compact_group_sum(const u64 *__restrict__ pfx_a, const u64 *__restrict__ pfx_b,
                                                    const i128 *__restrict__ base_a, const i128 *__restrict__ base_b, u32 ub,
                                                    u32 lb, u32 fast_cmax, double scale, int e0) {
    const u64 pa = pfx_a[ub], pb = pfx_b[lb];
    if (ub - lb < fast_cmax) return (double)(i64)(pa - pb) * scale;
    const i128 ba = base_a[ub >> GRB_BASE_SHIFT], bb = base_b[lb >> GRB_BASE_SHIFT];
    const i128 PA = ba + (i128)(i64)(pa - (u64)(u128)ba);
    const i128 PB = bb + (i128)(i64)(pb - (u64)(u128)bb);
    return fx_to_double(PA - PB, e0);
}


}  // namespace

// pipe.cu: the persistent pipelined kernel over the tiles of `b`; `fast` is a grb_op code (single reduction), PIPE_MULTI or
// PIPE_UBLB.  valid_bits != nullptr: validity goes out as one bit per result (bit i of word i / 32) instead of o.out_valid bytes
// (single-reduction calls only).
int grb_launch_pipe(grb_ctx *ctx, const Batch &b, const u32 *ls, const u32 *le, const TileOut &o, int fast, u32 *valid_bits);
