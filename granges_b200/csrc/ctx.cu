// ctx.cu -- context, error reporting, host/device buffer staging.
#include <stdarg.h>
#include <stdlib.h>

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
    GRB_CUDA(ctx, cudaMemcpyAsync(tmp.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
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
    if (host && bytes) GRB_CUDA(ctx, cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
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
    ctx->sweep_mode = 1;
    const char *sm = getenv("GRB_SWEEP_MODE");
    if (sm) ctx->sweep_mode = atoi(sm);
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
    if (cudaHostAlloc((void **)&ctx->stats_host, 1024 * 4, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&ctx->stats_dev, ctx->stats_host, 0) != cudaSuccess) {
        cudaGetLastError();
        ctx->stats_host = ctx->stats_dev = nullptr;  // order detection on device buffers is simply off
    } else {
        memset(ctx->stats_host, 0, 1024 * 4);
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
