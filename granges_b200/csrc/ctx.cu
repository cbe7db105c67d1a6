// ctx.cu -- context, error reporting, host/device buffer staging.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

int grb_fail(grb_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

bool grb_is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool grb_is_pageable_ptr(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// ---- page-locked staging rings -------------------------------------------------------------------------------------
// One ring = 4 page-locked buffers of 8 MB with an event each.  Rings are pooled process-wide (allocating page-locked
// memory costs milliseconds) and borrowed for the duration of one copy; each host thread that copies borrows its own.
namespace {

constexpr size_t RING_CHUNK = 8u << 20;
constexpr int RING_SLOTS = 4;
constexpr size_t STAGE_MIN = 1u << 20;  // smaller copies: the driver's own staging is fine

struct PinRing {
    void *buf[RING_SLOTS];
    cudaEvent_t ev[RING_SLOTS];
    int dev;
};

std::mutex g_ring_mu;
std::vector<PinRing *> g_ring_pool[64];  // per device: an event can only be recorded on a stream of the device it was created on

PinRing *ring_borrow() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ring_mu);
        if (!g_ring_pool[dev].empty()) {
            PinRing *r = g_ring_pool[dev].back();
            g_ring_pool[dev].pop_back();
            return r;
        }
    }
    PinRing *r = new PinRing();
    r->dev = dev;
    for (int i = 0; i < RING_SLOTS; i++) {
        r->buf[i] = nullptr;
        r->ev[i] = nullptr;
    }
    for (int i = 0; i < RING_SLOTS; i++) {
        if (cudaHostAlloc(&r->buf[i], RING_CHUNK, cudaHostAllocPortable) != cudaSuccess ||
            cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < RING_SLOTS; j++) {
                if (r->buf[j]) cudaFreeHost(r->buf[j]);
                if (r->ev[j]) cudaEventDestroy(r->ev[j]);
            }
            delete r;
            return nullptr;
        }
    }
    return r;
}

void ring_return(PinRing *r) {
    std::lock_guard<std::mutex> lk(g_ring_mu);
    g_ring_pool[r->dev].push_back(r);
}

}  // namespace

int grb_h2d_staged(grb_ctx *ctx, cudaStream_t stream, void *dst_dev, const void *src_host, size_t bytes) {
    if (bytes == 0) return GRB_OK;
    PinRing *r = (bytes >= STAGE_MIN && grb_is_pageable_ptr(src_host)) ? ring_borrow() : nullptr;
    if (!r) {
        GRB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
        return GRB_OK;
    }
    // (events belong to the device that is current when they are recorded: the ring's are recorded on `stream`'s device)
    cudaError_t e = cudaSuccess;
    int used = 0;
    for (size_t off = 0; off < bytes && e == cudaSuccess; off += RING_CHUNK, used++) {
        const int i = used % RING_SLOTS;
        const size_t len = bytes - off < RING_CHUNK ? bytes - off : RING_CHUNK;
        if (used >= RING_SLOTS) e = cudaEventSynchronize(r->ev[i]);  // the DMA out of this slot has finished
        if (e != cudaSuccess) break;
        memcpy(r->buf[i], (const char *)src_host + off, len);
        e = cudaMemcpyAsync((char *)dst_dev + off, r->buf[i], len, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaEventRecord(r->ev[i], stream);
    }
    // the ring goes back to the pool only when the engine is done with it
    for (int i = 0; i < RING_SLOTS && i < used; i++) {
        cudaError_t e2 = cudaEventSynchronize(r->ev[i]);
        if (e == cudaSuccess) e = e2;
    }
    ring_return(r);
    if (e != cudaSuccess) return grb_fail(ctx, GRB_ERR_CUDA, "staged host-to-device copy: %s", cudaGetErrorString(e));
    return GRB_OK;
}

int grb_d2h_staged(grb_ctx *ctx, cudaStream_t stream, void *dst_host, const void *src_dev, size_t bytes) {
    if (bytes == 0) return GRB_OK;
    PinRing *r = (bytes >= STAGE_MIN && grb_is_pageable_ptr(dst_host)) ? ring_borrow() : nullptr;
    if (!r) {
        GRB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
        return GRB_OK;
    }
    cudaError_t e = cudaSuccess;
    const size_t nchunks = (bytes + RING_CHUNK - 1) / RING_CHUNK;
    auto drain = [&](size_t c) {  // chunk c: wait for its DMA, copy it out of the ring
        const int i = (int)(c % RING_SLOTS);
        const size_t off = c * RING_CHUNK, len = bytes - off < RING_CHUNK ? bytes - off : RING_CHUNK;
        cudaError_t e2 = cudaEventSynchronize(r->ev[i]);
        if (e2 == cudaSuccess) memcpy((char *)dst_host + off, r->buf[i], len);
        return e2;
    };
    for (size_t c = 0; c < nchunks && e == cudaSuccess; c++) {
        const int i = (int)(c % RING_SLOTS);
        const size_t off = c * RING_CHUNK, len = bytes - off < RING_CHUNK ? bytes - off : RING_CHUNK;
        if (c >= RING_SLOTS) e = drain(c - RING_SLOTS);  // frees slot i
        if (e != cudaSuccess) break;
        e = cudaMemcpyAsync(r->buf[i], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaEventRecord(r->ev[i], stream);
    }
    for (size_t c = nchunks > RING_SLOTS ? nchunks - RING_SLOTS : 0; c < nchunks; c++) {
        cudaError_t e2 = drain(c);
        if (e == cudaSuccess) e = e2;
    }
    ring_return(r);
    if (e != cudaSuccess) return grb_fail(ctx, GRB_ERR_CUDA, "staged device-to-host copy: %s", cudaGetErrorString(e));
    return GRB_OK;
}

int InBuf::stage(grb_ctx *ctx, const void *src, size_t bytes) {
    if (bytes == 0 || src == nullptr) {
        d = src;
        return GRB_OK;
    }
    if (grb_is_device_ptr(src)) {
        d = src;
        return GRB_OK;
    }
    GRB_TRY(tmp.alloc(ctx, bytes));
    GRB_TRY(grb_h2d_staged(ctx, ctx->stream, tmp.p, src, bytes));
    d = tmp.p;
    return GRB_OK;
}

int OutBuf::stage(grb_ctx *ctx, void *dst, size_t nbytes) {
    bytes = nbytes;
    host = nullptr;
    if (nbytes == 0 || dst == nullptr) {
        d = dst;
        return GRB_OK;
    }
    if (grb_is_device_ptr(dst)) {
        d = dst;
        return GRB_OK;
    }
    GRB_TRY(tmp.alloc(ctx, nbytes));
    d = tmp.p;
    host = dst;
    return GRB_OK;
}

int OutBuf::finish(grb_ctx *ctx) {
    if (host && bytes) GRB_TRY(grb_d2h_staged(ctx, ctx->stream, host, d, bytes));
    return GRB_OK;
}

__global__ void k_copy_words(u32 *__restrict__ dst, const u32 *__restrict__ src, u32 nwords) {
    for (u32 i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

int grb_read_small(grb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    if (bytes > 256 || !ctx->stats_host || ctx->ctl_slot < 0 || ctx->ctl_slot > 15) {
        GRB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return GRB_OK;
    }
    const size_t off = 1024 + 64 * (size_t)ctx->ctl_slot;
    k_copy_words<<<1, 64, 0, ctx->stream>>>(ctx->stats_dev + off, (const u32 *)src_dev, (u32)((bytes + 3) / 4));
    GRB_LAUNCHED(ctx);
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(dst_host, ctx->stats_host + off, bytes);
    return GRB_OK;
}

int grb_check_flags(grb_ctx *ctx) {
    u32 f = 0;
    GRB_CUDA(ctx, cudaMemcpyAsync(&f, ctx->dev_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f) {
        GRB_CUDA(ctx, cudaMemsetAsync(ctx->dev_flags, 0, 4, ctx->stream));
        if (f & 1)
            return grb_fail(ctx, GRB_ERR_INVALID_RANGE,
                            "a left range has end <= start or end > 2^31-1 (the reference panics: ranges/mod.rs:30, "
                            "ranges/coitrees.rs:53-54); its result rows are those of an empty group");
    }
    return GRB_OK;
}

extern "C" {

int grb_abi_version(void) { return GRB_ABI_VERSION; }

int grb_ctx_create(int device, grb_ctx **out) {
    if (!out) return GRB_ERR_INVALID_ARG;
    *out = nullptr;
    grb_ctx *ctx = (grb_ctx *)calloc(1, sizeof(grb_ctx));
    if (!ctx) return GRB_ERR_OOM;
    ctx->device = device;
    ctx->tile_mode = 1;
    const char *tm = getenv("GRB_TILE_MODE");
    if (tm) ctx->tile_mode = atoi(tm);
    ctx->pipe_mode = 1;
    const char *pm = getenv("GRB_PIPE_MODE");
    if (pm) ctx->pipe_mode = atoi(pm);
    ctx->pipe_assign = 0;
    if (const char *pa = getenv("GRB_PIPE_ASSIGN")) ctx->pipe_assign = atoi(pa);
    ctx->sweep_mode = 1;
    const char *sm = getenv("GRB_SWEEP_MODE");
    if (sm) ctx->sweep_mode = atoi(sm);
    ctx->wave_ratio = 8;
    ctx->planar_rows = 1;
    if (const char *pr = getenv("GRB_PLANAR_ROWS")) ctx->planar_rows = atoi(pr);
    if (const char *wr = getenv("GRB_WAVE_RATIO")) ctx->wave_ratio = atoi(wr);
    // the remaining tuning knobs are read once, here, not on every call
    const char *up = getenv("GRB_UBLB_PIPE");
    ctx->ublb_pipe = up ? (atoi(up) != 0) : -1;
    const char *mm = getenv("GRB_MINMAX_MODE");
    ctx->minmax_mode = mm ? atoi(mm) : 1;
    const char *wc = getenv("GRB_WHOLE_CHUNKS");
    ctx->whole_chunks = wc && atoi(wc) > 0 ? atoi(wc) : 8;
    const char *ww = getenv("GRB_WHOLE_WORKERS");
    ctx->whole_workers = ww && atoi(ww) > 0 ? atoi(ww) : 4;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        // No CPU fallback: the caller gets a hard error.
        fprintf(stderr, "granges_b200: no usable CUDA device %d (%s)\n", device,
                e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
        free(ctx);
        return GRB_ERR_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        free(ctx);
        return GRB_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaDeviceGetDefaultMemPool(&ctx->pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;  // keep freed temporaries cached in the pool
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (cudaMalloc(&ctx->zeros, 2048) != cudaSuccess || cudaMemset(ctx->zeros, 0, 1024) != cudaSuccess ||
        cudaMemset((char *)ctx->zeros + 1024, 0xff, 1024) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        free(ctx);
        return GRB_ERR_OOM;
    }
    if (cudaMalloc((void **)&ctx->dev_flags, 16) != cudaSuccess || cudaMemset(ctx->dev_flags, 0, 16) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        free(ctx);
        return GRB_ERR_OOM;
    }
    // mapped pinned words: kernels report to the host without any stream operation (see grb_ctx::stats_host)
    if (cudaHostAlloc((void **)&ctx->stats_host, 2048 * 4, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&ctx->stats_dev, ctx->stats_host, 0) != cudaSuccess) {
        cudaGetLastError();
        ctx->stats_host = ctx->stats_dev = nullptr;  // order detection on device buffers is simply off
    } else {
        memset(ctx->stats_host, 0, 2048 * 4);
    }
    const char *lo = getenv("GRB_LEFT_ORDER");
    ctx->left_order = lo ? atoi(lo) : GRB_LEFT_ORDER_AUTO;
    *out = ctx;
    return GRB_OK;
}

int grb_ctx_set_left_order(grb_ctx *ctx, int order) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    if (order != GRB_LEFT_ORDER_AUTO && order != GRB_LEFT_ORDER_SORTED && order != GRB_LEFT_ORDER_UNSORTED)
        return grb_fail(ctx, GRB_ERR_INVALID_ARG, "set_left_order: unknown order %d", order);
    ctx->left_order = order;
    ctx->hist_ls = nullptr;
    return GRB_OK;
}

int grb_ctx_set_stream(grb_ctx *ctx, void *cuda_stream) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;  // a null handle is CUDA's default stream, as everywhere in CUDA
    return GRB_OK;
}

int grb_ctx_use_own_stream(grb_ctx *ctx) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    ctx->stream = ctx->own_stream;
    return GRB_OK;
}

int grb_ctx_sync(grb_ctx *ctx) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return grb_check_flags(ctx);
}

void grb_ctx_destroy(grb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->own_stream);
    cudaFree(ctx->zeros);
    cudaFree(ctx->dev_flags);
    cudaFree(ctx->batch_dev);
    if (ctx->stats_host) cudaFreeHost(ctx->stats_host);
    free(ctx->batch_host);
    free(ctx);
}

const char *grb_last_error(const grb_ctx *ctx) { return ctx ? ctx->err : "no context"; }

uint64_t grb_ctx_launch_count(const grb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int grb_ctx_sm_count(const grb_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
