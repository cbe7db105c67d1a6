// bucket.cu -- left ranges in any order.
//
// The reference answers lefts in whatever order they were pushed (GRanges::left_overlaps walks the left container,
// src/granges.rs:935-984) at one speed.  The tile kernels here are built for lefts whose neighbours are close in
// coordinate (sorted BED files, from_windows): a tile of 1024 such lefts touches a slice of the index that fits shared
// memory.  Shuffled lefts break that -- every left becomes a chain of dependent global loads, 20x slower.
//
// So unsorted lefts are first dropped into coordinate buckets of ~64 lefts per seqname (a counting sort: histogram with
// global atomics, one scan over the bucket counts, scatter through per-bucket cursors).  Order inside a bucket does not
// matter: a tile of bucketed lefts spans 1024 + 64 lefts' worth of coordinates, so its windows fit, and every result is
// written back to the row its left came from (`rank`).  Blocks walk the lefts in array order, i.e. seqname by seqname,
// and one seqname's share of every array involved fits the L2, so the atomics and the scattered 4- and 8-byte writes
// are absorbed there: DRAM sees each array about once.
#include <vector>

#include "common.cuh"

namespace {

struct BucketSeq {
    u64 left_begin, left_end;
    u32 bucket_base;  // first bucket of this seqname in the global bucket array
    u32 nb;           // buckets of this seqname
    int shift;        // bucket = min(l.start >> shift, nb - 1)
};

constexpr int BK_TPB = 256;
constexpr int BK_MAXSEQ = 1024;

__device__ __forceinline__ int bucket_seq_of(const BucketSeq *__restrict__ seqs, int nseq, u64 g) {
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seqs[mid].left_begin <= g)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// counts[bucket]++, and (for free) the number of adjacent inversions: how unsorted the lefts really are
__global__ void __launch_bounds__(BK_TPB) k_bucket_hist(const u32 *__restrict__ ls, const BucketSeq *__restrict__ seqs, int nseq, u64 first, u64 total,
                                                        u32 *__restrict__ counts, u32 *__restrict__ inversions) {
    const u64 g = first + (u64)blockIdx.x * BK_TPB + threadIdx.x;
    bool inv = false;
    if (g < total) {
        const BucketSeq q = seqs[bucket_seq_of(seqs, nseq, g)];
        const u32 x = ls[g];
        const u32 b = min(x >> q.shift, q.nb - 1u);
        atomicAdd(&counts[q.bucket_base + b], 1u);
        inv = g + 1 < q.left_end && ls[g + 1] < x;
    }
    const u32 m = __ballot_sync(GRB_FULL, inv);
    if (lane_id() == 0 && m) atomicAdd(inversions, (u32)__popc(m));
}

__global__ void __launch_bounds__(BK_TPB) k_bucket_scatter(const u32 *__restrict__ ls, const u32 *__restrict__ le, const BucketSeq *__restrict__ seqs,
                                                           int nseq, u64 first, u64 total, u32 *__restrict__ cursor, u32 *__restrict__ ls_b,
                                                           u32 *__restrict__ le_b, u32 *__restrict__ rank) {
    const u64 g = first + (u64)blockIdx.x * BK_TPB + threadIdx.x;
    if (g >= total) return;
    const BucketSeq q = seqs[bucket_seq_of(seqs, nseq, g)];
    const u32 x = ls[g];
    const u32 b = min(x >> q.shift, q.nb - 1u);
    const u64 p = first + atomicAdd(&cursor[q.bucket_base + b], 1u);  // cursor: exclusive scan of the counts, relative to `first`
    ls_b[p] = x;
    le_b[p] = le[g];
    rank[p] = (u32)(g - first);
}

// row i of the bucketed results goes back to the row its left came from
template <int K>
__global__ void __launch_bounds__(BK_TPB) k_unbucket(const double *__restrict__ out_b, const u8 *__restrict__ val_b, const u32 *__restrict__ rank, u64 first,
                                                     u64 total, int k, double *__restrict__ out, u8 *__restrict__ val) {
    const u64 i = first + (u64)blockIdx.x * BK_TPB + threadIdx.x;
    if (i >= total) return;
    const u64 r = first + rank[i];
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

// Drop the lefts [left_offsets[0], left_offsets[nseq]) of a batch into coordinate buckets.  extent[s] = a coordinate at or
// beyond which seqname s holds no right (lefts past it share the last bucket).  ls_b / le_b are indexed like ls / le
// (absolute positions, so tiles stay aligned), and so is rank: rank[p] = original row (relative to the batch's first left) of the left now at p.  inversions_dev (may be NULL)
// receives the number of adjacent inversions seen -- the caller's feedback on whether bucketing was needed at all.
int grb_bucket_lefts(grb_ctx *ctx, const u64 *left_offsets, const u64 *extent, int nseq, const u32 *ls, const u32 *le, u32 *ls_b, u32 *le_b,
                     u32 *rank, u32 *inversions_dev) {
    if (nseq > BK_MAXSEQ * 64) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "bucket: too many seqnames in one call");
    const u64 first = left_offsets[0], total = left_offsets[nseq];
    if (total - first >= 0xffffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "bucket: more than 2^32-1 lefts in one call");
    std::vector<BucketSeq> h((size_t)nseq);
    u64 nbuckets = 0;
    for (int s = 0; s < nseq; s++) {
        BucketSeq q;
        q.left_begin = left_offsets[s];
        q.left_end = left_offsets[s + 1];
        const u64 nl = q.left_end - q.left_begin;
        u64 want = nl / 64 + 1;  // ~64 lefts per bucket: a tile of 1024 bucketed lefts spans 1024 + 64 lefts' worth of coordinates
        if (want > (1u << 22)) want = 1u << 22;
        const u64 ext = extent[s] ? extent[s] : 1;
        int shift = 0;
        while (shift < 31 && (ext >> shift) + 1 > want) shift++;
        q.shift = shift;
        q.nb = (u32)((ext >> shift) + 1);
        q.bucket_base = (u32)nbuckets;
        nbuckets += q.nb;
        h[s] = q;
    }
    if (nbuckets >= 0x7fffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "bucket: too many buckets");
    TempBuf d_seqs, counts, cursor, inv_tmp;
    GRB_TRY(d_seqs.alloc(ctx, sizeof(BucketSeq) * nseq));
    GRB_TRY(counts.alloc(ctx, nbuckets * 4));
    GRB_TRY(cursor.alloc(ctx, nbuckets * 4));
    if (!inversions_dev) {
        GRB_TRY(inv_tmp.alloc(ctx, 4));
        inversions_dev = inv_tmp.as<u32>();
    }
    // (pageable source: the copy is staged before the call returns, so the vector may go out of scope)
    GRB_CUDA(ctx, cudaMemcpyAsync(d_seqs.p, h.data(), sizeof(BucketSeq) * nseq, cudaMemcpyHostToDevice, ctx->stream));
    GRB_CUDA(ctx, cudaMemsetAsync(counts.p, 0, nbuckets * 4, ctx->stream));
    GRB_CUDA(ctx, cudaMemsetAsync(inversions_dev, 0, 4, ctx->stream));
    const unsigned blocks = (unsigned)((total - first + BK_TPB - 1) / BK_TPB);
    if (blocks == 0) return GRB_OK;
    k_bucket_hist<<<blocks, BK_TPB, 0, ctx->stream>>>(ls, d_seqs.as<BucketSeq>(), nseq, first, total, counts.as<u32>(), inversions_dev);
    GRB_LAUNCHED(ctx);
    GRB_TRY(grb_scan_u32_to_u32(ctx, counts.as<u32>(), cursor.as<u32>(), nbuckets, false));
    k_bucket_scatter<<<blocks, BK_TPB, 0, ctx->stream>>>(ls, le, d_seqs.as<BucketSeq>(), nseq, first, total, cursor.as<u32>(), ls_b, le_b, rank);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

int grb_unbucket_results(grb_ctx *ctx, u64 first, u64 total, int k, const double *out_b, const u8 *val_b, const u32 *rank, double *out, u8 *val) {
    const unsigned blocks = (unsigned)((total - first + BK_TPB - 1) / BK_TPB);
    if (blocks == 0) return GRB_OK;
    if (k == 1)
        k_unbucket<1><<<blocks, BK_TPB, 0, ctx->stream>>>(out_b, val_b, rank, first, total, k, out, val);
    else
        k_unbucket<0><<<blocks, BK_TPB, 0, ctx->stream>>>(out_b, val_b, rank, first, total, k, out, val);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}
