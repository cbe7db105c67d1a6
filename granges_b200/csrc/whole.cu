// whole.cu -- `granges map` in one call for host-resident inputs (src/commands.rs:477-547: into_coitrees -> map_data ->
// left_overlaps -> map_joins), pipelined per seqname so that the PCIe link never idles, and the same call spread over
// the GPUs of one box.
//
//   worker threads (private streams)   seqname s: lefts + rights + scores -> device, index build   (H2D engine + SMs)
//                                      later: results of finished seqnames -> host                 (D2H engine)
//   caller's thread                    per chunk of seqnames: wait for their indexes, map kernels, hand the downloads out
//
// The two-call path (grb_index_build_batch, then grb_map_joins_f64_batch) uploads ALL rights before the first result can
// travel back; here the downloads of chunk c overlap the uploads of the later chunks.  Results are those of the two-call
// path bit for bit (same kernels, same indexes).
//
// Host buffers may be pageable (a Rust Vec, a numpy array): every copy goes through grb_h2d_staged / grb_d2h_staged,
// i.e. through page-locked staging rings filled / drained by the worker threads themselves, so the memcpy work into and
// out of the rings is spread over all of them.
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

#include <chrono>

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Job {
    int kind;  // 0: upload + index seqname s;  1: download rows [a, b) once event `after` has fired;  -1: stop
    int s;
    u64 a, b;
    cudaEvent_t after;
};

}  // namespace

extern "C" int grb_map_joins_whole_f64(grb_ctx *ctx, int nseq, const uint32_t *const *r_starts, const uint32_t *const *r_ends,
                                       const size_t *r_n, const double *const *r_scores, const uint8_t *const *r_valid,
                                       const uint64_t *left_offsets, const uint32_t *ls, const uint32_t *le, const int *ops, int k,
                                       double *out, uint8_t *out_valid) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    if (nseq <= 0 || !r_starts || !r_ends || !r_n || !left_offsets || !ops)
        return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: NULL argument");
    if (k < 1 || k > GRB_MAX_OPS) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: k must be in [1, %d]", GRB_MAX_OPS);
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    const u64 first = left_offsets[0], total = left_offsets[nseq];
    if (total > first && (!ls || !le || !out || !out_valid)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: NULL buffer");
    for (int s = 0; s < nseq; s++)
        if (left_offsets[s + 1] < left_offsets[s]) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_whole: left_offsets must be non-decreasing");
    if (total == first) return GRB_OK;
    const bool timing = getenv("GRB_TIMING") != nullptr;  // debug: per-job wall times on stderr
    const double t_call = now_ms();
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // earlier work on the caller's stream precedes the workers'
    // The lefts are on the host: sample their order here, once, and tell the per-chunk calls (which only see device copies)
    const int saved_order = ctx->left_order;
    if (saved_order == GRB_LEFT_ORDER_AUTO) {
        int disorder = 0;
        GRB_TRY(grb_left_disorder(ctx, ls, first, total, &disorder));
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
        // The tail of the call is what follows the last upload: the last chunk's index builds, its map pass and the download of
        // its rows.  Let the last seqnames be chunks of their own, so that only ONE seqname's results are still on the device
        // when the uploads end (with three seqnames in the last chunk the call ended 2 ms later: tools/e2e_timeline.py).
        const int tail = nseq < 4 ? nseq : 4;
        std::vector<int> fine;
        for (int c : cuts)
            if (c < nseq - tail) fine.push_back(c);
        for (int s = nseq - tail; s <= nseq; s++) fine.push_back(s);
        if (fine.empty() || fine.front() != 0) fine.insert(fine.begin(), 0);
        cuts.swap(fine);
    }
    const int nchunks = (int)cuts.size() - 1;

    // ---- device copies of the lefts and of the results (indexed like the caller's arrays: tiles stay aligned)
    TempBuf d_ls, d_le, d_out, d_val;
    GRB_TRY(d_ls.alloc(ctx, total * 4));
    GRB_TRY(d_le.alloc(ctx, total * 4));
    GRB_TRY(d_out.alloc(ctx, total * (u64)k * 8));
    GRB_TRY(d_val.alloc(ctx, total * (u64)k));
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the buffers exist before any worker stream touches them

    // ---- workers
    std::vector<grb_index *> idx((size_t)nseq, nullptr);
    std::vector<std::atomic<int>> built((size_t)nseq);
    for (auto &b : built) b.store(0);
    std::mutex mu;
    std::condition_variable cv_built, cv_jobs;
    std::deque<Job> jobs;
    std::atomic<int> failed{0};
    // pageable buffers keep every worker busy with memcpy: use more of them
    const bool pageable = grb_is_pageable_ptr(ls) || grb_is_pageable_ptr(out) || (r_starts[0] && grb_is_pageable_ptr(r_starts[0]));
    int wmax = ctx->whole_workers;  // GRB_WHOLE_WORKERS
    if (pageable && wmax < 8) wmax = 8;
    const int nworkers = nseq < wmax ? (nseq < 2 ? 2 : nseq) : wmax;
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
            return grb_fail(ctx, GRB_ERR_CUDA, "map_joins_whole: cannot create a stream");
        }
        clones[w].stream = clones[w].own_stream;
    }
    for (int s = 0; s < nseq; s++) jobs.push_back(Job{0, s, 0, 0, nullptr});
    auto work = [&](int w) {
        grb_ctx *c = &clones[w];
        cudaSetDevice(c->device);
        while (true) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_jobs.wait(lk, [&] { return !jobs.empty(); });
                j = jobs.front();
                if (j.kind < 0) return;  // (the stop job stays in the queue for the other workers)
                jobs.pop_front();
            }
            int rc = GRB_OK;
            const double t_job = now_ms();
            double t_up = 0, t_ix = 0;
            if (j.kind == 0) {
                const int s = j.s;
                if (!failed.load()) {
                    const u64 a = left_offsets[s], b = left_offsets[s + 1];
                    if (b > a) {
                        rc = grb_h2d_staged(c, c->stream, d_ls.as<u32>() + a, ls + a, (b - a) * 4);
                        if (rc == GRB_OK) rc = grb_h2d_staged(c, c->stream, d_le.as<u32>() + a, le + a, (b - a) * 4);
                    }
                    if (timing) {
                        cudaStreamSynchronize(c->stream);
                        t_up = now_ms();
                    }
                    if (rc == GRB_OK && r_n[s]) {
                        // the scores start travelling before the index build: its kernels and its host round trips then run
                        // beside a copy that is already queued, and the link does not idle while this worker waits for them
                        TempBuf d_sc, d_va;
                        const bool have_sc = r_scores && r_scores[s];
                        const bool have_va = have_sc && r_valid && r_valid[s];
                        if (have_sc) {
                            rc = d_sc.alloc(c, r_n[s] * 8);
                            if (rc == GRB_OK) rc = grb_h2d_staged(c, c->stream, d_sc.p, r_scores[s], r_n[s] * 8);
                            if (rc == GRB_OK && have_va) rc = d_va.alloc(c, r_n[s]);
                            if (rc == GRB_OK && have_va) rc = grb_h2d_staged(c, c->stream, d_va.p, r_valid[s], r_n[s]);
                        }
                        if (rc == GRB_OK) rc = grb_index_build(c, r_starts[s], r_ends[s], nullptr, r_n[s], &idx[s]);
                        t_ix = now_ms();
                        if (rc == GRB_OK && have_sc)
                            rc = grb_index_set_scores(idx[s], d_sc.as<double>(), have_va ? d_va.as<u8>() : nullptr, r_n[s]);
                        if (rc == GRB_OK) idx[s]->ctx = ctx;  // the clone goes away; the index is used on the caller's context
                    }
                    if (rc == GRB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = grb_fail(c, GRB_ERR_CUDA, "map_joins_whole: upload failed");
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    built[s].store(1);
                }
                cv_built.notify_all();
            } else if (!failed.load()) {
                cudaStreamWaitEvent(c->stream, j.after, 0);
                rc = grb_d2h_staged(c, c->stream, out + j.a * (u64)k, d_out.as<double>() + j.a * (u64)k, (j.b - j.a) * (u64)k * 8);
                if (rc == GRB_OK) rc = grb_d2h_staged(c, c->stream, out_valid + j.a * (u64)k, d_val.as<u8>() + j.a * (u64)k, (j.b - j.a) * (u64)k);
            }
            if (timing)
                fprintf(stderr, "[whole dev %d] worker %d %s seq %d: %.2f .. %.2f ms (lefts up %.2f, index built %.2f)\n", ctx->device, w,
                        j.kind == 0 ? "upload+index" : "download", j.s, t_job - t_call, now_ms() - t_call, t_up ? t_up - t_call : 0.0,
                        t_ix ? t_ix - t_call : 0.0);
            if (rc != GRB_OK) {
                rcs[w] = rc;
                failed.store(1);
                cv_built.notify_all();
            }
        }
    };
    std::vector<std::thread> threads;
    for (int w = 0; w < nworkers; w++) threads.emplace_back(work, w);
    cv_jobs.notify_all();

    // ---- caller's thread: map kernels chunk by chunk, downloads handed to the workers
    int rc = GRB_OK;
    std::vector<cudaEvent_t> done((size_t)nchunks, nullptr);
    for (int c = 0; c < nchunks; c++) cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming);
    for (int c = 0; c < nchunks && rc == GRB_OK; c++) {
        const int s0 = cuts[c], s1 = cuts[c + 1];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_built.wait(lk, [&] {
                if (failed.load()) return true;
                for (int s = s0; s < s1; s++)
                    if (!built[s].load()) return false;
                return true;
            });
        }
        if (failed.load()) break;
        // absolute offsets + whole-array pointers: tiles stay aligned to multiples of 1024 lefts
        rc = grb_map_joins_f64_batch(ctx, idx.data() + s0, left_offsets + s0, s1 - s0, d_ls.as<u32>(), d_le.as<u32>(), ops, k, d_out.as<double>(),
                                     d_val.as<u8>());
        cudaEventRecord(done[c], ctx->stream);
        if (timing) fprintf(stderr, "[whole dev %d] map chunk %d (seq %d..%d) enqueued at %.2f ms\n", ctx->device, c, s0, s1, now_ms() - t_call);
        if (rc == GRB_OK) {
            std::lock_guard<std::mutex> lk(mu);
            // downloads go to the FRONT of the queue: the next free worker starts them, so results travel back while the
            // later seqnames are still being uploaded and indexed (PCIe is full duplex)
            for (int s = s1 - 1; s >= s0; s--)
                if (left_offsets[s + 1] > left_offsets[s]) jobs.push_front(Job{1, s, left_offsets[s], left_offsets[s + 1], done[c]});
        }
        cv_jobs.notify_all();
    }
    if (rc != GRB_OK) failed.store(1);
    {
        std::lock_guard<std::mutex> lk(mu);
        jobs.push_back(Job{-1, 0, 0, 0, nullptr});
    }
    cv_jobs.notify_all();
    for (auto &t : threads) t.join();
    cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
    for (int w = 0; w < nworkers; w++) {
        cudaError_t e = cudaStreamSynchronize(clones[w].own_stream);
        if (e != cudaSuccess && e3 == cudaSuccess) e3 = e;
        cudaStreamDestroy(clones[w].own_stream);
        ctx->launches += clones[w].launches;
        if (rcs[w] != GRB_OK && rc == GRB_OK) {
            rc = rcs[w];
            memcpy(ctx->err, clones[w].err, sizeof(ctx->err));
        }
    }
    for (int c = 0; c < nchunks; c++)
        if (done[c]) cudaEventDestroy(done[c]);
    for (int s = 0; s < nseq; s++)
        if (idx[s]) {
            idx[s]->ctx = ctx;
            grb_index_destroy(idx[s]);
        }
    if (timing) fprintf(stderr, "[whole dev %d] done at %.2f ms\n", ctx->device, now_ms() - t_call);
    if (rc != GRB_OK) return rc;
    if (e3 != cudaSuccess) return grb_fail(ctx, GRB_ERR_CUDA, "map_joins_whole: %s", cudaGetErrorString(e3));
    return grb_check_flags(ctx);
}

// ---- the same call over several GPUs of one box ---------------------------------------------------------------------
//
// Overlaps never cross seqnames (src/granges.rs:958-962), so whole seqnames are independent units: the seqnames are cut
// into ndev contiguous runs of near-equal transfer volume (a linear partition -- contiguous, so that every GPU's rows are
// one contiguous block of the caller's output arrays), one host thread per GPU runs grb_map_joins_whole_f64 on its run,
// and every run writes its own rows of the caller's `out` / `out_valid`: that IS the host-side concatenation in the
// reference's row order (src/iterators.rs:43-76).  No data moves between GPUs.

namespace {

std::mutex g_mgpu_mu;
grb_ctx *g_mgpu_ctx[64] = {};

}  // namespace

extern "C" int grb_map_joins_whole_f64_mgpu(const int *devices, int ndev, int nseq, const uint32_t *const *r_starts,
                                            const uint32_t *const *r_ends, const size_t *r_n, const double *const *r_scores,
                                            const uint8_t *const *r_valid, const uint64_t *left_offsets, const uint32_t *ls, const uint32_t *le,
                                            const int *ops, int k, double *out, uint8_t *out_valid, char *err, size_t err_len) {
    auto fail = [&](int code, const char *msg) {
        if (err && err_len) snprintf(err, err_len, "%s", msg);
        return code;
    };
    if (err && err_len) err[0] = 0;
    if (!devices || ndev < 1 || ndev > 64 || nseq <= 0 || !r_starts || !r_ends || !r_n || !left_offsets || !ops)
        return fail(GRB_ERR_INVALID_ARG, "map_joins_whole_mgpu: bad argument");
    for (int d = 0; d < ndev; d++)
        if (devices[d] < 0 || devices[d] >= 64) return fail(GRB_ERR_INVALID_ARG, "map_joins_whole_mgpu: device index out of range");
    std::lock_guard<std::mutex> call_lock(g_mgpu_mu);  // one multi-GPU call at a time: the per-device contexts are shared
    // ---- linear partition of the seqnames by transfer volume (what bounds a GPU's share: PCIe)
    std::vector<u64> cost((size_t)nseq);
    for (int s = 0; s < nseq; s++) cost[s] = (left_offsets[s + 1] - left_offsets[s]) * (8 + 9ull * k) + (u64)r_n[s] * 16 + 1;
    const int parts = ndev < nseq ? ndev : nseq;
    // best[p][i] = smallest possible maximum over p runs covering seqnames [0, i)
    std::vector<std::vector<u64>> best((size_t)parts + 1, std::vector<u64>((size_t)nseq + 1, ~0ull));
    std::vector<std::vector<int>> from((size_t)parts + 1, std::vector<int>((size_t)nseq + 1, 0));
    std::vector<u64> pre((size_t)nseq + 1, 0);
    for (int s = 0; s < nseq; s++) pre[s + 1] = pre[s] + cost[s];
    best[0][0] = 0;
    for (int p = 1; p <= parts; p++)
        for (int i = p; i <= nseq; i++)
            for (int j = p - 1; j < i; j++) {
                if (best[p - 1][j] == ~0ull) continue;
                const u64 run = pre[i] - pre[j];
                const u64 v = best[p - 1][j] > run ? best[p - 1][j] : run;
                if (v < best[p][i]) {
                    best[p][i] = v;
                    from[p][i] = j;
                }
            }
    std::vector<int> cut((size_t)parts + 1);
    cut[parts] = nseq;
    for (int p = parts; p >= 1; p--) cut[p - 1] = from[p][cut[p]];
    // ---- one context per device, kept for later calls
    for (int p = 0; p < parts; p++) {
        const int dev = devices[p];
        if (!g_mgpu_ctx[dev]) {
            int rc = grb_ctx_create(dev, &g_mgpu_ctx[dev]);
            if (rc != GRB_OK) return fail(rc, "map_joins_whole_mgpu: cannot create a context on one of the devices");
        }
    }
    std::vector<int> rcs((size_t)parts, GRB_OK);
    std::vector<std::thread> threads;
    for (int p = 0; p < parts; p++) {
        threads.emplace_back([&, p] {
            const int s0 = cut[p], s1 = cut[p + 1];
            if (s1 <= s0) return;
            grb_ctx *c = g_mgpu_ctx[devices[p]];
            rcs[p] = grb_map_joins_whole_f64(c, s1 - s0, r_starts + s0, r_ends + s0, r_n + s0, r_scores ? r_scores + s0 : nullptr,
                                             r_valid ? r_valid + s0 : nullptr, left_offsets + s0, ls, le, ops, k, out, out_valid);
        });
    }
    for (auto &t : threads) t.join();
    for (int p = 0; p < parts; p++)
        if (rcs[p] != GRB_OK) return fail(rcs[p], grb_last_error(g_mgpu_ctx[devices[p]]));
    return GRB_OK;
}
