// windows.cu -- GRangesEmpty::from_windows (src/granges.rs:634-666) in closed form, on the device.
//
// The reference loop, per seqname of length len (step defaults to width):
//     start = 0; while start < len { end = start + width;
//         if end >= len { if chop { break } else { push (start, min(end, len)) } }   // and keeps stepping
//         else { push (start, end) }
//         start += step }
// so without chop every start k*step < len yields a window (truncated ones included -- pinned by
// src/granges.rs:1876-1907), and with chop exactly the starts with start + width < len do.
#include <stdlib.h>

#include <vector>

#include "common.cuh"

struct grb_ranges {
    grb_ctx *ctx;
    u64 n;
    int nseq;
    u64 *seq_offsets;  // host, nseq + 1
    u32 *seq_id, *starts, *ends;  // device
};

namespace {

struct WinSeq {
    u64 off;
    u32 len;
    u32 pad;
};

__global__ void k_windows(const WinSeq *__restrict__ seqs, int nseq, u64 total, u32 width, u32 step, u32 *__restrict__ seq_id,
                          u32 *__restrict__ starts, u32 *__restrict__ ends) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (seqs[mid].off <= i)
            lo = mid;
        else
            hi = mid;
    }
    u64 k = i - seqs[lo].off;
    u64 st = k * (u64)step;
    u64 en = st + width;
    u64 len = seqs[lo].len;
    seq_id[i] = (u32)lo;
    starts[i] = (u32)st;
    ends[i] = (u32)(en < len ? en : len);
}

}  // namespace

extern "C" {

int grb_windows(grb_ctx *ctx, const uint32_t *seqlens, int nseq, uint32_t width, uint32_t step, int chop,
                grb_ranges **out) {
    if (!ctx || !out) return GRB_ERR_INVALID_ARG;
    *out = nullptr;
    if (nseq < 0 || (nseq && !seqlens)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "windows: NULL seqlens");
    if (width == 0) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "windows: width must be > 0");
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    const u64 st = step ? step : width;
    grb_ranges *r = (grb_ranges *)calloc(1, sizeof(grb_ranges));
    if (!r) return grb_fail(ctx, GRB_ERR_OOM, "windows: host allocation failed");
    r->ctx = ctx;
    r->nseq = nseq;
    r->seq_offsets = (u64 *)calloc((size_t)nseq + 1, sizeof(u64));
    std::vector<WinSeq> ws((size_t)nseq + 1);
    u64 total = 0;
    for (int s = 0; s < nseq; s++) {
        u64 len = seqlens[s], cnt;
        if (chop)
            cnt = len > width ? (len - width - 1) / st + 1 : 0;  // starts with start + width < len
        else
            cnt = (len + st - 1) / st;  // starts k*step < len
        ws[s].off = total;
        ws[s].len = seqlens[s];
        r->seq_offsets[s] = total;
        total += cnt;
    }
    ws[nseq].off = total;
    r->seq_offsets[nseq] = total;
    r->n = total;
    int rc = GRB_OK;
    do {
        if (grb_dev_malloc(ctx, (void **)&r->seq_id, total * 4 + 16) != cudaSuccess || grb_dev_malloc(ctx, (void **)&r->starts, total * 4 + 16) != cudaSuccess ||
            grb_dev_malloc(ctx, (void **)&r->ends, total * 4 + 16) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_OOM, "windows: cannot allocate %llu windows on the device", (unsigned long long)total);
            break;
        }
        if (total == 0) break;
        TempBuf d;
        if ((rc = d.alloc(ctx, sizeof(WinSeq) * (nseq + 1)))) break;
        cudaMemcpyAsync(d.p, ws.data(), sizeof(WinSeq) * (nseq + 1), cudaMemcpyHostToDevice, ctx->stream);
        u64 blocks = (total + 255) / 256;
        k_windows<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d.as<WinSeq>(), nseq, total, width, (u32)st, r->seq_id, r->starts, r->ends);
        GRB_LAUNCHED_BREAK(ctx, rc)
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = grb_fail(ctx, GRB_ERR_CUDA, "windows: %s", cudaGetErrorString(e));
    } while (0);
    if (rc != GRB_OK) {
        grb_ranges_destroy(r);
        return rc;
    }
    *out = r;
    return GRB_OK;
}

size_t grb_ranges_len(const grb_ranges *r) { return r ? (size_t)r->n : 0; }

int grb_ranges_seq_offsets(const grb_ranges *r, uint64_t *offsets) {
    if (!r || !offsets) return GRB_ERR_INVALID_ARG;
    for (int s = 0; s <= r->nseq; s++) offsets[s] = r->seq_offsets[s];
    return GRB_OK;
}

int grb_ranges_copy(const grb_ranges *r, uint32_t *seq_id, uint32_t *starts, uint32_t *ends) {
    if (!r) return GRB_ERR_INVALID_ARG;
    grb_ctx *ctx = r->ctx;
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (r->n) {
        if (seq_id) GRB_CUDA(ctx, cudaMemcpyAsync(seq_id, r->seq_id, r->n * 4, cudaMemcpyDefault, ctx->stream));
        if (starts) GRB_CUDA(ctx, cudaMemcpyAsync(starts, r->starts, r->n * 4, cudaMemcpyDefault, ctx->stream));
        if (ends) GRB_CUDA(ctx, cudaMemcpyAsync(ends, r->ends, r->n * 4, cudaMemcpyDefault, ctx->stream));
    }
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

const uint32_t *grb_ranges_dev_starts(const grb_ranges *r) { return r ? r->starts : nullptr; }
const uint32_t *grb_ranges_dev_ends(const grb_ranges *r) { return r ? r->ends : nullptr; }

void grb_ranges_destroy(grb_ranges *r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    grb_dev_free(r->ctx, r->seq_id);
    grb_dev_free(r->ctx, r->starts);
    grb_dev_free(r->ctx, r->ends);
    free(r->seq_offsets);
    free(r);
}

}  // extern "C"
