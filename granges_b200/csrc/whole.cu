// whole.cu -- `granges map` in one call for host-resident inputs (src/commands.rs:477-547: into_coitrees -> map_data ->
// left_overlaps -> map_joins), pipelined per seqname so that the PCIe link never idles:
//
//   worker threads (private streams)   rights + scores of seqname s -> device, index build          (H2D engine + SMs)
//   caller's thread                    lefts of chunk c -> device | map kernels | results -> host   (H2D | SMs | D2H engine)
//
// The two-call path (grb_index_build_batch, then grb_map_joins_f64_batch) uploads ALL rights before the first result can
// travel back, so the downloads only overlap the lefts' upload; here the downloads of chunk c overlap the uploads of the
// later chunks' rights too.  Results are those of the two-call path bit for bit (same kernels, same indexes).
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" int grb_map_joins_whole_f64(grb_ctx *ctx, int nseq, const uint32_t *const *r_starts, const uint32_t *const *r_ends,
                                       const size_t *r_n, const double *const *r_scores, const uint8_t *const *r_valid,
                                       const uint64_t *left_offsets, const uint32_t *ls, const uint32_t *le, const int *ops, int k,
                                       double *out, uint8_t *out_valid) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    if (nseq <= 0 || !r_starts || !r_ends || !r_n || !left_offsets || !ops)
        return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: NULL argument");
    if (k < 1 || k > GRB_MAX_OPS) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: k must be in [1, %d]", GRB_MAX_OPS);
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    const u64 total = left_offsets[nseq];
    if (total && (!ls || !le || !out || !out_valid)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: NULL buffer");
    for (int s = 0; s < nseq; s++)
        if (left_offsets[s + 1] < left_offsets[s]) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: left_offsets must be non-decreasing");
    if (total == 0) return GRB_OK;
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // earlier work on the caller's stream precedes the workers'
    // The lefts are on the host: sample their order here, once, and tell the per-chunk calls (which only see device copies)
    const int saved_order = ctx->left_order;
    if (saved_order == GRB_LEFT_ORDER_AUTO) {
        int disorder = 0;
        GRB_TRY(grb_left_disorder(ctx, ls, 0, total, &disorder));
        ctx->left_order = disorder > 100 ? GRB_LEFT_ORDER_UNSORTED : GRB_LEFT_ORDER_SORTED;
    }
    struct RestoreOrder {
        grb_ctx *c;
        int v;
        ~RestoreOrder() { c->left_order = v; }
    } restore_order{ctx, saved_order};

    // ---- chunks of whole seqnames, about an eighth of the transfer volume each
    std::vector<int> cuts{0};
    {
        u64 vol_total = 0;
        for (int s = 0; s < nseq; s++) vol_total += (left_offsets[s + 1] - left_offsets[s]) * (8 + 9ull * k) + (u64)r_n[s] * 16;
        const u64 nch = (u64)ctx->whole_chunks;  // GRB_WHOLE_CHUNKS
        const u64 target = vol_total / nch + 1;
        u64 acc = 0;
        for (int s = 0; s < nseq; s++) {
            acc += (left_offsets[s + 1] - left_offsets[s]) * (8 + 9ull * k) + (u64)r_n[s] * 16;
            if (acc >= target && s + 1 < nseq) {
                cuts.push_back(s + 1);
                acc = 0;
            }
        }
        cuts.push_back(nseq);
    }
    const int nchunks = (int)cuts.size() - 1;

    // ---- index builders
    std::vector<grb_index *> idx((size_t)nseq, nullptr);
    std::vector<std::atomic<int>> built((size_t)nseq);
    for (auto &b : built) b.store(0);
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> failed{0};
    const int wmax = ctx->whole_workers;  // GRB_WHOLE_WORKERS
    const int nworkers = nseq < wmax ? nseq : wmax;
    std::vector<grb_ctx> clones((size_t)nworkers);
    std::vector<int> rcs((size_t)nworkers, GRB_OK);
    for (int w = 0; w < nworkers; w++) {
        clones[w] = *ctx;  // shares the device, the pool and the zero / flag words
        clones[w].batch_dev = nullptr;
        clones[w].batch_host = nullptr;
        clones[w].batch_bytes = clones[w].batch_cap = 0;
        clones[w].launches = 0;
        clones[w].err[0] = 0;
        if (cudaStreamCreateWithFlags(&clones[w].own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            for (int v = 0; v < w; v++) cudaStreamDestroy(clones[v].own_stream);
            return grb_fail(ctx, GRB_ERR_CUDA, "map_joins_whole: cannot create a stream");
        }
        clones[w].stream = clones[w].own_stream;
    }
    auto work = [&](int w) {
        grb_ctx *c = &clones[w];
        cudaSetDevice(c->device);
        for (int s = w; s < nseq && !failed.load(); s += nworkers) {
            int rc = GRB_OK;
            if (r_n[s]) {
                rc = grb_index_build(c, r_starts[s], r_ends[s], nullptr, r_n[s], &idx[s]);
                if (rc == GRB_OK && r_scores && r_scores[s])
                    rc = grb_index_set_scores(idx[s], r_scores[s], (r_valid && r_valid[s]) ? r_valid[s] : nullptr, r_n[s]);
                if (rc == GRB_OK) idx[s]->ctx = ctx;  // the clone goes away; the index is used on the caller's context
            }
            if (rc != GRB_OK) {
                rcs[w] = rc;
                failed.store(1);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                built[s].store(1);
            }
            cv.notify_all();
        }
        std::lock_guard<std::mutex> lk(mu);
        cv.notify_all();
    };
    std::vector<std::thread> threads;
    for (int w = 0; w < nworkers; w++) threads.emplace_back(work, w);

    // ---- caller's thread: lefts up, map, results down, chunk by chunk
    int rc = GRB_OK;
    TempBuf d_ls, d_le, d_out, d_val;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> up((size_t)nchunks, nullptr), done((size_t)nchunks, nullptr);
    cudaEvent_t ready = nullptr;
    do {
        if ((rc = d_ls.alloc(ctx, total * 4))) break;
        if ((rc = d_le.alloc(ctx, total * 4))) break;
        if ((rc = d_out.alloc(ctx, total * (u64)k * 8))) break;
        if ((rc = d_val.alloc(ctx, total * (u64)k))) break;
        if (cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "map_joins_whole: cannot create a stream");
            break;
        }
        cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
        cudaEventRecord(ready, ctx->stream);  // the staging buffers exist from here on
        cudaStreamWaitEvent(h2d, ready, 0);
        for (int c = 0; c < nchunks; c++) {
            cudaEventCreateWithFlags(&up[c], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming);
        }
        for (int c = 0; c < nchunks && rc == GRB_OK; c++) {
            const int s0 = cuts[c], s1 = cuts[c + 1];
            const u64 a = left_offsets[s0], b = left_offsets[s1];
            if (b > a) {
                cudaMemcpyAsync(d_ls.as<u32>() + a, ls + a, (b - a) * 4, cudaMemcpyHostToDevice, h2d);
                cudaMemcpyAsync(d_le.as<u32>() + a, le + a, (b - a) * 4, cudaMemcpyHostToDevice, h2d);
            }
            cudaEventRecord(up[c], h2d);
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    if (failed.load()) return true;
                    for (int s = s0; s < s1; s++)
                        if (!built[s].load()) return false;
                    return true;
                });
            }
            if (failed.load()) break;
            cudaStreamWaitEvent(ctx->stream, up[c], 0);
            // absolute offsets + whole-array pointers: tiles stay aligned to multiples of 1024 lefts
            rc = grb_map_joins_f64_batch(ctx, idx.data() + s0, left_offsets + s0, s1 - s0, d_ls.as<u32>(), d_le.as<u32>(), ops, k,
                                         d_out.as<double>(), d_val.as<u8>());
            cudaEventRecord(done[c], ctx->stream);
            cudaStreamWaitEvent(d2h, done[c], 0);
            if (b > a && rc == GRB_OK) {
                cudaMemcpyAsync(out + a * (u64)k, d_out.as<double>() + a * (u64)k, (b - a) * (u64)k * 8, cudaMemcpyDeviceToHost, d2h);
                cudaMemcpyAsync(out_valid + a * (u64)k, d_val.as<u8>() + a * (u64)k, (b - a) * (u64)k, cudaMemcpyDeviceToHost, d2h);
            }
        }
    } while (0);
    if (rc != GRB_OK) failed.store(1);
    cv.notify_all();
    for (auto &t : threads) t.join();
    cudaError_t e1 = h2d ? cudaStreamSynchronize(h2d) : cudaSuccess, e2 = d2h ? cudaStreamSynchronize(d2h) : cudaSuccess,
                e3 = cudaStreamSynchronize(ctx->stream);
    for (int c = 0; c < nchunks; c++) {
        if (up[c]) cudaEventDestroy(up[c]);
        if (done[c]) cudaEventDestroy(done[c]);
    }
    if (ready) cudaEventDestroy(ready);
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
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
        if (idx[s]) {
            idx[s]->ctx = ctx;
            grb_index_destroy(idx[s]);
        }
    if (rc != GRB_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return grb_fail(ctx, GRB_ERR_CUDA, "map_joins_whole: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
    return grb_check_flags(ctx);
}
