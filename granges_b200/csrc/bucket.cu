// bucket.cu -- left ranges in any order.
//
// The reference answers lefts in whatever order they were pushed (GRanges::left_overlaps walks the left container,
// src/granges.rs:935-984) at one speed.  The tile kernels here are built for lefts whose neighbours are close in
// coordinate (sorted BED files, from_windows): a tile of 1024 such lefts touches a slice of the index that fits shared
// memory.  Shuffled lefts break that -- every left becomes a chain of dependent global loads, 20x slower.
//
// So unsorted lefts are first dropped into coordinate buckets of ~64 lefts per seqname, in two levels of 256-way
// partitions that only use shared-memory atomics: level A sends the lefts of every tile of 4096 to the seqname's 256
// coarse buckets (per-tile histograms -> one scan -> scatter), level B splits every coarse bucket (a few thousand lefts,
// contiguous by then, L2-resident) into its 256 fine buckets with one CTA.  Order inside a bucket does not matter: a
// tile of bucketed lefts spans 1024 + 64 lefts' worth of coordinates, so its windows fit the pipeline's ring slots.
// Level B also records where every original row went (`pos`), and the results are gathered back through it: sequential
// writes, gathers out of a region the L2 holds.  (A first version with global atomics per left -- histogram and cursor
// bumps -- spent 5.4 ms of an 8.2 ms pass at 100M lefts on 40 G atomics/s: profiles/r2/unsorted_lefts.txt.)
#include <vector>

#include "common.cuh"

namespace {

struct BucketSeq {
    u64 left_begin, left_end;
    u32 tile_begin;   // first tile of this seqname (tiles of BK_TILE lefts, never across seqnames)
    u32 ntiles;
    u32 hist_base;    // first entry of this seqname in the (digit-major) histogram: 256 * tile_begin
    u32 nb;           // fine buckets of this seqname (<= 65536)
    int shift;        // fine bucket f = min(l.start >> shift, nb - 1); coarse digit = f >> 8, fine digit = f & 255
};

constexpr int BK_TPB = 256;
#ifndef BK_ITEMS
#define BK_ITEMS 16  // lefts per thread of a level-A tile (4096-left tiles; 16384-left tiles measured slower: 9.8 vs 9.1 ms)
#endif
#ifndef BK_FINE_TPB
#define BK_FINE_TPB 1024  // threads of a level-B CTA: fewer, larger CTAs keep the buckets being split inside the L2
#endif
constexpr int BK_TILE = BK_TPB * BK_ITEMS;

__device__ __forceinline__ int seq_of_tile(const BucketSeq *__restrict__ seqs, int nseq, u32 t) {
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seqs[mid].tile_begin <= t)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Level A, pass 1: per tile, how many lefts fall into each of the seqname's 256 coarse buckets (shared-memory atomics
// only); and, for free, the number of adjacent inversions: how unsorted the lefts really are.
__global__ void __launch_bounds__(BK_TPB) k_coarse_hist(const u32 *__restrict__ ls, const BucketSeq *__restrict__ seqs, int nseq,
                                                        u32 *__restrict__ hist, u32 *__restrict__ inversions) {
    __shared__ u32 h[256];
    __shared__ u32 s_inv;
    h[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_inv = 0;
    __syncthreads();
    const BucketSeq q = seqs[seq_of_tile(seqs, nseq, blockIdx.x)];
    const u32 tis = blockIdx.x - q.tile_begin;
    const u64 base = q.left_begin + (u64)tis * BK_TILE;
    u32 inv = 0;
#pragma unroll 8
    for (int i = 0; i < BK_ITEMS; i++) {
        const u64 g = base + (u64)i * BK_TPB + threadIdx.x;
        if (g < q.left_end) {
            const u32 x = ls[g];
            atomicAdd(&h[min(x >> q.shift, q.nb - 1u) >> 8], 1u);
            inv += (g + 1 < q.left_end && ls[g + 1] < x) ? 1u : 0u;
        }
    }
    inv = warp_sum(inv);
    if (lane_id() == 0 && inv) atomicAdd(&s_inv, inv);
    __syncthreads();
    hist[q.hist_base + threadIdx.x * q.ntiles + tis] = h[threadIdx.x];
    if (threadIdx.x == 0 && s_inv) atomicAdd(inversions, s_inv);
}

// Both partition levels stage a tile of 4096 lefts in shared memory grouped by bucket before anything is written, so that
// consecutive threads write consecutive addresses of a bucket's run (16 lefts = 64 bytes per run and array on average):
// written straight from the threads that read them, every left costs one 32-byte transaction per array -- 300M of them
// at 100M lefts, which is what bounded the first version (2.5 + 2.7 ms for the two levels).
struct BucketStage {
    u32 ls[BK_TILE];
    u32 le[BK_TILE];
    u32 rank[BK_TILE];
    u8 bkt[BK_TILE];
    u32 lstart[256];   // first staged slot of each bucket
    u32 cursor[256];   // next free staged slot of each bucket
    u32 gbase[256];    // where the bucket's run of this tile starts in global memory
    u32 wsum[8];
};

// exclusive scan of 256 values held one per thread by threads 0..255 (the first 8 warps); returns the exclusive prefix
__device__ __forceinline__ u32 bk_scan256(u32 v, u32 *wsum) {
    u32 incl = v;
    if (threadIdx.x < 256) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(GRB_FULL, incl, o);
            if (lane_id() >= (u32)o) incl += t;
        }
        if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    }
    __syncthreads();
    u32 woff = 0;
    if (threadIdx.x < 256)
        for (u32 w = 0; w < (threadIdx.x >> 5); w++) woff += wsum[w];
    return woff + incl - v;
}

// Level A, pass 2: the lefts of a tile go to their coarse buckets; offs = exclusive scan of hist (positions relative to
// the batch's first left), so the tile's own counts are hist itself.  Order inside a bucket is whatever the shared-memory
// atomics give: it does not matter.
__global__ void __launch_bounds__(BK_TPB) k_coarse_scatter(const u32 *__restrict__ ls, const u32 *__restrict__ le, const BucketSeq *__restrict__ seqs,
                                                           int nseq, u64 first, const u32 *__restrict__ hist, const u32 *__restrict__ offs,
                                                           u32 *__restrict__ ls_a, u32 *__restrict__ le_a, u32 *__restrict__ rank_a) {
    extern __shared__ __align__(16) unsigned char bk_raw[];
    BucketStage &S = *reinterpret_cast<BucketStage *>(bk_raw);
    const BucketSeq q = seqs[seq_of_tile(seqs, nseq, blockIdx.x)];
    const u32 tis = blockIdx.x - q.tile_begin;
    const u32 hidx = q.hist_base + threadIdx.x * q.ntiles + tis;
    const u32 mine = hist[hidx];
    S.gbase[threadIdx.x] = offs[hidx];
    const u32 ex = bk_scan256(mine, S.wsum);
    S.lstart[threadIdx.x] = ex;
    S.cursor[threadIdx.x] = ex;
    __syncthreads();
    const u64 base = q.left_begin + (u64)tis * BK_TILE;
    const u32 nhere = (u32)(q.left_end - base < (u64)BK_TILE ? q.left_end - base : (u64)BK_TILE);
#pragma unroll 4
    for (int i = 0; i < BK_ITEMS; i++) {
        const u32 o = (u32)i * BK_TPB + threadIdx.x;
        if (o < nhere) {
            const u32 x = ls[base + o];
            const u32 b = min(x >> q.shift, q.nb - 1u) >> 8;
            const u32 slot = atomicAdd(&S.cursor[b], 1u);
            S.ls[slot] = x;
            S.le[slot] = le[base + o];
            S.rank[slot] = (u32)(base + o - first);
            S.bkt[slot] = (u8)b;
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < nhere; i += BK_TPB) {
        const u32 b = S.bkt[i];
        const u32 p = S.gbase[b] + (i - S.lstart[b]);
        ls_a[p] = S.ls[i];
        le_a[p] = S.le[i];
        rank_a[p] = S.rank[i];
    }
}

// Level B: one CTA per coarse bucket (its lefts are contiguous after level A) splits it into its 256 fine buckets -- a
// count over the whole bucket, then tile after tile staged by fine bucket and written as runs -- and records where every
// original row went.
__global__ void __launch_bounds__(BK_FINE_TPB) k_fine_partition(const u32 *__restrict__ ls_a, const u32 *__restrict__ le_a, const u32 *__restrict__ rank_a,
                                                           const BucketSeq *__restrict__ seqs, u64 first, const u32 *__restrict__ offs,
                                                           u32 *__restrict__ ls_b, u32 *__restrict__ le_b, u32 *__restrict__ pos) {
    extern __shared__ __align__(16) unsigned char bk_raw[];
    BucketStage &S = *reinterpret_cast<BucketStage *>(bk_raw);
    __shared__ u32 cnt[256];
    const BucketSeq q = seqs[blockIdx.x >> 8];
    const u32 nthr = blockDim.x;
    const u32 c = blockIdx.x & 255u;
    if (q.ntiles == 0) return;
    const u32 b0 = offs[q.hist_base + c * q.ntiles], b1 = offs[q.hist_base + (c + 1) * q.ntiles];  // (the scan has one entry past the end)
    if (b1 <= b0) return;
    if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = b0 + threadIdx.x; i < b1; i += nthr) atomicAdd(&cnt[min(ls_a[i] >> q.shift, q.nb - 1u) & 255u], 1u);
    __syncthreads();
    {
        const u32 v = threadIdx.x < 256 ? cnt[threadIdx.x] : 0u;
        const u32 ex = bk_scan256(v, S.wsum);
        if (threadIdx.x < 256) S.gbase[threadIdx.x] = b0 + ex;  // next global position of every fine bucket
    }
    __syncthreads();
    for (u32 t0 = b0; t0 < b1; t0 += BK_TILE) {
        const u32 nhere = b1 - t0 < (u32)BK_TILE ? b1 - t0 : (u32)BK_TILE;
        if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
        __syncthreads();
        // the tile's own counts; every left remembers its place inside its bucket
        constexpr int PER = BK_TILE / BK_FINE_TPB;
        u32 x[PER], lr[PER];
#pragma unroll
        for (int e = 0; e < PER; e++) {
            const u32 o = (u32)e * nthr + threadIdx.x;
            x[e] = 0;
            lr[e] = 0;
            if (o < nhere) {
                x[e] = ls_a[t0 + o];
                lr[e] = atomicAdd(&cnt[min(x[e] >> q.shift, q.nb - 1u) & 255u], 1u);
            }
        }
        __syncthreads();
        {
            const u32 v = threadIdx.x < 256 ? cnt[threadIdx.x] : 0u;
            const u32 ex = bk_scan256(v, S.wsum);
            if (threadIdx.x < 256) S.lstart[threadIdx.x] = ex;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < PER; e++) {
            const u32 o = (u32)e * nthr + threadIdx.x;
            if (o < nhere) {
                const u32 b = min(x[e] >> q.shift, q.nb - 1u) & 255u;
                const u32 slot = S.lstart[b] + lr[e];
                S.ls[slot] = x[e];
                S.le[slot] = le_a[t0 + o];
                S.rank[slot] = rank_a[t0 + o];
                S.bkt[slot] = (u8)b;
            }
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < nhere; i += nthr) {
            const u32 b = S.bkt[i];
            const u32 p = S.gbase[b] + (i - S.lstart[b]);
            ls_b[first + p] = S.ls[i];
            le_b[first + p] = S.le[i];
            pos[S.rank[i]] = p;
        }
        __syncthreads();
        if (threadIdx.x < 256) S.gbase[threadIdx.x] += cnt[threadIdx.x];
        __syncthreads();
    }
}

// row r of the results = row pos[r] of the bucketed results: sequential writes, gathers out of a region the L2 holds
template <int K>
__global__ void __launch_bounds__(BK_TPB) k_unbucket(const double *__restrict__ out_b, const u8 *__restrict__ val_b, const u32 *__restrict__ pos, u64 first,
                                                     u64 total, int k, double *__restrict__ out, u8 *__restrict__ val) {
    const u64 r = first + (u64)blockIdx.x * BK_TPB + threadIdx.x;
    if (r >= total) return;
    const u64 i = first + pos[r - first];
    if (K == 1) {
        out[r] = out_b[i];
        val[r] = val_b[i];
    } else {
        for (int c = 0; c < k; c++) {
            out[r * (u64)k + c] = out_b[i * (u64)k + c];
            val[r * (u64)k + c] = val_b[i * (u64)k + c];
        }
    }
}

__global__ void k_sample_inversions(const u32 *__restrict__ ls, u64 first, u64 nsamp, u64 stride, u32 *__restrict__ result) {
    // thread t looks at the adjacent pair at first + t * stride (seqname boundaries count as ordinary pairs: a handful)
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 g = first + t * stride;
    const bool inv = t < nsamp && ls[g + 1] < ls[g];
    const u32 m = __ballot_sync(GRB_FULL, inv);
    if (lane_id() == 0 && m) atomicAdd(result, (u32)__popc(m));
}

}  // namespace

// Sampled share of adjacent left pairs that are out of order, in 1/1024ths (0 = sorted, ~512 = shuffled).  Host buffers
// are sampled on the host; device buffers by a tiny kernel and a 4-byte read-back (a synchronisation).
int grb_left_disorder(grb_ctx *ctx, const u32 *ls, u64 first, u64 total, int *per_1024) {
    *per_1024 = 0;
    const u64 n = total - first;
    if (n < 2) return GRB_OK;
    const u64 nsamp = n - 1 < 4096 ? n - 1 : 4096;
    const u64 stride = (n - 1) / nsamp;
    u32 inv = 0;
    if (!grb_is_device_ptr(ls)) {
        for (u64 t = 0; t < nsamp; t++) {
            const u64 g = first + t * stride;
            inv += ls[g + 1] < ls[g];
        }
    } else {
        TempBuf res;
        GRB_TRY(res.alloc(ctx, 4));
        GRB_CUDA(ctx, cudaMemsetAsync(res.p, 0, 4, ctx->stream));
        k_sample_inversions<<<(unsigned)((nsamp + 127) / 128), 128, 0, ctx->stream>>>(ls, first, nsamp, stride, res.as<u32>());
        GRB_LAUNCHED(ctx);
        GRB_CUDA(ctx, cudaMemcpyAsync(&inv, res.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *per_1024 = (int)((u64)inv * 1024 / nsamp);
    return GRB_OK;
}

// Drop the lefts [left_offsets[0], left_offsets[nseq]) of a batch into coordinate buckets of ~64 lefts.  extent[s] = a
// coordinate at or beyond which seqname s holds no right (lefts past it share the last bucket).  ls_b / le_b are indexed
// like ls / le (absolute positions, so tiles stay aligned); pos[r] = where row r (relative to the batch's first left) went
// (relative likewise).  inversions_dev (may be NULL) receives the number of adjacent inversions seen -- the caller's
// feedback on whether bucketing was needed at all.
int grb_bucket_lefts(grb_ctx *ctx, const u64 *left_offsets, const u64 *extent, int nseq, const u32 *ls, const u32 *le, u32 *ls_b, u32 *le_b,
                     u32 *pos, u32 *inversions_dev) {
    const u64 first = left_offsets[0], total = left_offsets[nseq];
    const u64 n = total - first;
    if (n >= 0xffffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "bucket: more than 2^32-1 lefts in one call");
    if ((u64)nseq * 256 >= 0x7fffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "bucket: too many seqnames in one call");
    std::vector<BucketSeq> h((size_t)nseq);
    u64 tiles = 0;
    for (int s = 0; s < nseq; s++) {
        BucketSeq q;
        q.left_begin = left_offsets[s];
        q.left_end = left_offsets[s + 1];
        const u64 nl = q.left_end - q.left_begin;
        u64 want = nl / 64 + 1;  // ~64 lefts per bucket: a tile of 1024 bucketed lefts spans 1024 + 64 lefts' worth of coordinates
        if (want > 65536) want = 65536;
        const u64 ext = extent[s] ? extent[s] : 1;
        int shift = 0;
        while (shift < 31 && (ext >> shift) + 1 > want) shift++;
        q.shift = shift;
        q.nb = (u32)((ext >> shift) + 1);
        q.tile_begin = (u32)tiles;
        q.ntiles = (u32)((nl + BK_TILE - 1) / BK_TILE);
        q.hist_base = (u32)(tiles * 256);
        tiles += q.ntiles;
        h[s] = q;
    }
    if (tiles * 256 >= 0x7fffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "bucket: too many lefts in one call");
    if (tiles == 0) return GRB_OK;
    const size_t nh = (size_t)tiles * 256;
    TempBuf d_seqs, hist, offs, inv_tmp, ls_a, le_a, rank_a;
    GRB_TRY(d_seqs.alloc(ctx, sizeof(BucketSeq) * nseq));
    GRB_TRY(hist.alloc(ctx, nh * 4));
    GRB_TRY(offs.alloc(ctx, (nh + 1) * 4));
    GRB_TRY(ls_a.alloc(ctx, n * 4));
    GRB_TRY(le_a.alloc(ctx, n * 4));
    GRB_TRY(rank_a.alloc(ctx, n * 4));
    if (!inversions_dev) {
        GRB_TRY(inv_tmp.alloc(ctx, 4));
        inversions_dev = inv_tmp.as<u32>();
    }
    // (pageable source: the copy is staged before the call returns, so the vector may go out of scope)
    GRB_CUDA(ctx, cudaMemcpyAsync(d_seqs.p, h.data(), sizeof(BucketSeq) * nseq, cudaMemcpyHostToDevice, ctx->stream));
    GRB_CUDA(ctx, cudaMemsetAsync(inversions_dev, 0, 4, ctx->stream));
    k_coarse_hist<<<(unsigned)tiles, BK_TPB, 0, ctx->stream>>>(ls, d_seqs.as<BucketSeq>(), nseq, hist.as<u32>(), inversions_dev);
    GRB_LAUNCHED(ctx);
    GRB_TRY(grb_scan_u32_to_u32(ctx, hist.as<u32>(), offs.as<u32>(), nh, true));
    {
        // the opt-in to > 48 KB of dynamic shared memory is per device
        static bool configured[64] = {};
        if (ctx->device < 0 || ctx->device >= 64 || !configured[ctx->device]) {
            GRB_CUDA(ctx, cudaFuncSetAttribute(k_coarse_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketStage)));
            GRB_CUDA(ctx, cudaFuncSetAttribute(k_fine_partition, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketStage)));
            if (ctx->device >= 0 && ctx->device < 64) configured[ctx->device] = true;
        }
    }
    k_coarse_scatter<<<(unsigned)tiles, BK_TPB, sizeof(BucketStage), ctx->stream>>>(ls, le, d_seqs.as<BucketSeq>(), nseq, first, hist.as<u32>(),
                                                                                   offs.as<u32>(), ls_a.as<u32>(), le_a.as<u32>(), rank_a.as<u32>());
    GRB_LAUNCHED(ctx);
    k_fine_partition<<<(unsigned)nseq * 256u, BK_FINE_TPB, sizeof(BucketStage), ctx->stream>>>(ls_a.as<u32>(), le_a.as<u32>(), rank_a.as<u32>(),
                                                                                         d_seqs.as<BucketSeq>(), first, offs.as<u32>(), ls_b, le_b, pos);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

int grb_unbucket_results(grb_ctx *ctx, u64 first, u64 total, int k, const double *out_b, const u8 *val_b, const u32 *pos, double *out, u8 *val) {
    const unsigned blocks = (unsigned)((total - first + BK_TPB - 1) / BK_TPB);
    if (blocks == 0) return GRB_OK;
    if (k == 1)
        k_unbucket<1><<<blocks, BK_TPB, 0, ctx->stream>>>(out_b, val_b, pos, first, total, k, out, val);
    else
        k_unbucket<0><<<blocks, BK_TPB, 0, ctx->stream>>>(out_b, val_b, pos, first, total, k, out, val);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}
