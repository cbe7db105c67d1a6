// common.cuh -- shared types, error plumbing and device helpers of libgranges_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "granges_b200.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned __int128 u128;
typedef __int128 i128;

#define GRB_FULL 0xffffffffu

// ---- context -------------------------------------------------------------

struct grb_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;
    cudaMemPool_t pool;
    u64 launches;
    // device copy of the last batch's per-seqname descriptors (reused when the next call has the same batch)
    void *batch_dev;
    void *batch_host;
    size_t batch_bytes, batch_cap;
    u32 *dev_flags; // device word: bit 0 = a left range with end <= start or end > 2^31-1 was seen by a kernel
    void *zeros;    // 1 KB of zeros (tables) + 1 KB of 0xff (sentinel keys): what an absent seqname / empty index reads
    int pipe_assign; // GRB_PIPE_ASSIGN (A/B): 1 = runs of consecutive tiles per CTA instead of round-robin single tiles
    int pipe_mode;  // 1 = persistent warp-specialised pipeline for single-reduction maps (default), 0 = k_tile only (debug)
    int tile_mode;  // 1 = TMA-staged smem windows (default), 0 = global binary search only (debug)
    int ublb_pipe;  // GRB_UBLB_PIPE: 1 / 0 = the (ub, pp, lb) pass always / never in the pipelined kernel; -1 (default) = beside sums
    int minmax_mode; // GRB_MINMAX_MODE: 0 = MIN / MAX sweep the members even without a median beside them (debug)
    int whole_chunks, whole_workers;  // GRB_WHOLE_CHUNKS / GRB_WHOLE_WORKERS (defaults 8 / 4)
    int left_order; // grb_left_order: how the order of the lefts is decided (grb_ctx_set_left_order)
    // Mapped pinned words the kernels report to without any stream operation (device alias: stats_dev).  Every report carries
    // the generation number of the call that asked for it, written last, so the host never mistakes an older report for news:
    //   [0] adjacent inversions seen by a bucket pass, [2] its generation;
    //   [16 + b] tiles CTA b of a pipelined launch could not stage, [1] that launch's generation (written by CTA 0).
    u32 *stats_host;
    u32 *stats_dev;
    int ctl_slot;   // which 256-byte slot of the control area (words 1024.. of the mapped block) grb_read_small uses: 0, or 1 + worker for clones
    const void *hist_ls;  // AUTO on device buffers: the lefts of the previous call ...
    u64 hist_first, hist_total;
    u32 hist_tiles;
    int hist_bucketed;    // ... the verdict in force for them (kept until a newer report says otherwise) ...
    u32 hist_gen0;        // ... and the first generation that belongs to them
    u32 gen;              // generation counter of the reporting calls
    int planar_rows; // GRB_PLANAR_ROWS (default 1): beside k_sweep3 the prefix-sum columns travel through a side array and the sweep writes whole rows
    int wave_ratio; // GRB_WAVE_RATIO: medians of a run of lefts come from per-run wavelet trees when lefts * ratio >= slice rights (default 8, 0 = never)
    int sweep_mode; // 1 = staged rank sweep k_sweep3 for MIN / MAX / MEDIAN (default), 2 = k_sweep3 without staging, 0 = k_sweep2 only (debug)
    char err[512];
};

int grb_fail(grb_ctx *ctx, int code, const char *fmt, ...);

#define GRB_CUDA(ctx, expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return grb_fail((ctx), _e == cudaErrorMemoryAllocation ? GRB_ERR_OOM : GRB_ERR_CUDA,    \
                            "%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
    } while (0)

#define GRB_TRY(...)           \
    do {                       \
        int _rc = (__VA_ARGS__); \
        if (_rc != GRB_OK) return _rc; \
    } while (0)

// Count + check one kernel launch.
#define GRB_LAUNCHED(ctx)                                                                     \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return grb_fail((ctx), GRB_ERR_CUDA, "%s:%d: kernel launch: %s", __FILE__, __LINE__, \
                            cudaGetErrorString(_e));                                          \
    } while (0)

// The same inside a `do { ... } while (0)` body that reports through a local `rc` and `break`.
#define GRB_LAUNCHED_BREAK(ctx, rc)                                                                   \
    (ctx)->launches++;                                                                                \
    if (cudaError_t _e = cudaGetLastError(); _e != cudaSuccess) {                                     \
        rc = grb_fail((ctx), GRB_ERR_CUDA, "%s:%d: kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
        break;                                                                                        \
    }

// Stream-ordered temporary device buffer (pool keeps the memory cached, no sync on free).
struct TempBuf {
    grb_ctx *ctx = nullptr;
    void *p = nullptr;
    size_t bytes = 0;
    TempBuf() {}
    TempBuf(const TempBuf &) = delete;
    TempBuf &operator=(const TempBuf &) = delete;
    int alloc(grb_ctx *c, size_t nbytes) {
        release();
        ctx = c;
        bytes = nbytes ? nbytes : 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, c->stream);
        if (e != cudaSuccess) {
            p = nullptr;
            return grb_fail(c, e == cudaErrorMemoryAllocation ? GRB_ERR_OOM : GRB_ERR_CUDA,
                            "cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
        }
        return GRB_OK;
    }
    void release() {
        if (p) cudaFreeAsync(p, ctx->stream);
        p = nullptr;
    }
    template <typename T>
    T *as() const {
        return (T *)p;
    }
    ~TempBuf() { release(); }
};

// An input that may live on the host or on the device: dev() is always a device pointer.
struct InBuf {
    TempBuf tmp;
    const void *d = nullptr;
    int stage(grb_ctx *ctx, const void *src, size_t bytes);
    template <typename T>
    const T *as() const {
        return (const T *)d;
    }
};

// An output that may live on the host or on the device: kernels write dev(); finish() copies back.
struct OutBuf {
    TempBuf tmp;
    void *d = nullptr;
    void *host = nullptr;
    size_t bytes = 0;
    int stage(grb_ctx *ctx, void *dst, size_t nbytes);
    int finish(grb_ctx *ctx);  // enqueue D2H if the destination is a host buffer
    template <typename T>
    T *as() const {
        return (T *)d;
    }
};

bool grb_is_device_ptr(const void *p);
// Is p ordinary pageable host memory (a Rust Vec, a malloc'ed buffer, a numpy array) -- neither device memory nor
// page-locked?  cudaMemcpyAsync on such memory is staged by the driver in small synchronous pieces.
bool grb_is_pageable_ptr(const void *p);
// Host <-> device copies on `stream` that go through the library's own ring of page-locked staging buffers when the host
// side is pageable and the copy is large: the calling thread copies chunk c + 1 into the ring while the DMA engine moves
// chunk c.  Page-locked host memory takes the plain asynchronous copy.  grb_d2h_staged returns when the data is in `dst`
// (pageable) or when the copy is enqueued (page-locked).
int grb_h2d_staged(grb_ctx *ctx, cudaStream_t stream, void *dst_dev, const void *src_host, size_t bytes);
int grb_d2h_staged(grb_ctx *ctx, cudaStream_t stream, void *dst_host, const void *src_dev, size_t bytes);
// A few words of device memory -> host, now: a one-warp kernel stores them into the context's mapped page-locked block and the
// stream is synchronised.  Not a cudaMemcpy: a small device-to-host copy queues behind every bulk download that is in flight
// on the copy engine (results of earlier seqnames travelling back), and the index build that waits for it stalls for milliseconds.
int grb_read_small(grb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
// Read-and-clear the device flag word after a synchronisation; turns a flagged invalid left range into
// GRB_ERR_INVALID_RANGE (the reference cannot even construct such a range: src/ranges/mod.rs:30, coitrees.rs:53-54).
int grb_check_flags(grb_ctx *ctx);

// Long-lived device arrays (indexes, pairs, windows) also come from the stream-ordered pool: plain
// cudaMalloc / cudaFree cost milliseconds each once a process holds many GB, the pool costs microseconds.
inline cudaError_t grb_dev_malloc(grb_ctx *ctx, void **p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream); }
inline void grb_dev_free(grb_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

// ---- index ----------------------------------------------------------------

// How SUM / MEAN are answered for one index.
//   SWEEP    the members are summed one by one (subnormal scores, or a dynamic range beyond 126 bits)
//   COMPACT  W + 8 <= 62: prefixes mod 2^64 per right + one exact i128 base per 256 rights
//   WIDE     one exact i128 prefix per right (W too large for WIDE96)
//   WIDE96   W + 8 <= 94: prefixes mod 2^96 per right as two planes (u64 low, u32 high) + the per-256 bases
// Non-finite scores do not change the mode: they count as 0 in the fixed point and are tallied by the nf_* prefix
// counters, from which the IEEE result of the left-to-right sum (NaN / +inf / -inf) follows by rule.
enum { GRB_SUM_SWEEP = 0, GRB_SUM_COMPACT = 1, GRB_SUM_WIDE = 2, GRB_SUM_WIDE96 = 3 };

#define GRB_BASE_SHIFT 8  // compact mode: one exact i128 base per 256 rights beside the 64-bit wrapped prefixes
#define GRB_BME_SHIFT 6   // block-max-end granularity: 64 rights

struct grb_index {
    grb_ctx *ctx;
    u32 n;
    // start order: sorted by (start, end, data index)
    u32 *starts;  // n + pad, padded with 0xffffffff
    u32 *ends;    // n + pad
    u32 *didx;    // n
    u32 *bme;     // ceil(n/64): max end of each block of 64 rights (start order)
    // end order
    u32 *ends_sorted;  // n + pad, padded with 0xffffffff
    u32 *perm_e;       // n: position in start order of the i-th smallest end
    // scores (set_scores)
    bool has_scores, has_nulls, has_nan;
    u64 n_data;
    double *score_a;  // n, start order; None -> 0.0
    u32 *vbits_a;     // ceil(n/32) validity bits (start order), only if has_nulls
    u32 *vcnt_a;      // n+1 prefix count of valid scores (start order), only if has_nulls
    u32 *vcnt_b;      // n+1 prefix count of valid scores (end order), only if has_nulls
    int sum_mode;
    int e0;       // fixed-point unit = 2^e0
    int width;    // bits of the largest |score| in fixed point (W)
    u64 *pfx_a;   // compact: n+1 prefixes mod 2^64 (start order)
    u64 *pfx_b;   // compact: n+1 prefixes mod 2^64 (end order)
    u32 *pfh_a;   // wide96: bits 64..95 of the prefixes (start order); pfx_a holds bits 0..63
    u32 *pfh_b;
    uint4 *nf_a;  // only if has_nonfinite: n+1 prefix counts (x = +inf, y = -inf, z = NaN) in start order
    uint4 *nf_b;  // ... in end order
    u16 *vc16_a;  // only if has_nulls: vcnt_a & 0xffff (what the pipelined kernel stages)
    u16 *vc16_b;
    i128 *base_a; // compact: (n>>8)+1 exact block bases;  wide: n+1 exact prefixes
    i128 *base_b;
    // coordinate-bin lookup tables: lut[b] = #keys < (b << lut_shift), b in [0, lut_bmax + 1]
    int lut_shift;
    u32 lut_bmax;
    u32 *lut_a;   // over starts
    u32 *lut_b;   // over ends_sorted
    u16 *lut16_a;  // lut_a & 0xffff: within one staged window (< 65536 keys) the low halves are enough
    u16 *lut16_b;
    // member-sweep tables (MIN / MAX / MEDIAN), built on first use by grb_index_prepare_members
    bool has_nonfinite;
    bool members_ready;
    u32 n_valid;          // rights that carry a score
    u32 *rank_a;          // n + pad, start order: rank of the right's score among the valid scores, ordered by
                          // (score, position) -- all different; GRB_RANK_NONE when the score is None
    double *score_sorted; // n_valid: the scores in rank order
    // MIN / MAX tables, built on first use by grb_index_prepare_minmax (ranks as in rank_a)
    bool minmax_ready;
    u32 *stab_min;        // 2n + 1: smallest rank among the rights with start < x < end, indexed by #{start < x} + #{end <= x}
    int *stab_max;        // 2n + 1: largest such rank, -1 if none
    u32 *sp_min;          // (sp_levels) x sp_nb: sparse table of the per-32-block rank minima in start order
    int *sp_max;
    u32 sp_nb;
    int sp_levels;
    // overlap-bp tables, built on first use by grb_index_prepare_bp
    u64 *sum_starts;      // n + 1: sum of the i smallest starts
    u64 *sum_ends;        // n + 1: sum of the i smallest ends
    // weighted-mean tables, built on first use by grb_index_prepare_wm (exact-sum indexes only)
    i128 *wsum_a;         // n + 1: exact prefix of score * start (fixed point), start order
    i128 *wsum_b;         // n + 1: exact prefix of score * end, end order
    u64 *sum_starts_v;    // n + 1: like sum_starts over the rights that carry a score (only if has_nulls)
    u64 *sum_ends_v;
    bool wm_ready, wm_unavailable;
    u32 *fb;              // ceil(n/64): fb[b] <= position of the earliest right that straddles any coordinate
                          // >= starts[64 b]  (how far back a window over the rights must reach)
};

#define GRB_RANK_NONE 0xffffffffu

// CSR form of a left grouped join (grb_left_overlaps)
struct grb_pairs {
    grb_ctx *ctx;
    const grb_index *idx;
    u64 nl, np;
    u32 *ls, *le;   // device copies of the lefts
    u64 *offsets;   // device, nl + 1
    u32 *pos;       // device, np positions in the index's start order
};


// Per-seqname view handed to the kernels (one entry per seqname of a batch).
struct SeqDev {
    const u32 *starts, *ends, *ends_sorted, *didx, *bme;
    const double *score_a;
    const u32 *vbits_a, *vcnt_a, *vcnt_b;
    const u64 *pfx_a, *pfx_b;
    const u32 *pfh_a, *pfh_b;
    const uint4 *nf_a, *nf_b;
    const u16 *vc16_a, *vc16_b;
    const i128 *base_a, *base_b;
    const u32 *lut_a, *lut_b;
    const u16 *lut16_a, *lut16_b;
    const u32 *rank_a, *fb, *perm_e;
    const double *score_sorted;
    const u32 *stab_min, *sp_min;
    const int *stab_max, *sp_max;
    u32 sp_nb;
    const u64 *sum_starts, *sum_ends, *sum_starts_v, *sum_ends_v;
    const i128 *wsum_a, *wsum_b;
    u64 left_begin, left_end;
    double scale;    // 2^e0
    u32 n;
    u32 tile_begin;
    u32 lut_bmax;
    u32 fast_cmax;   // groups smaller than this have sums that fit 63 bits: wrapped 64-bit prefixes suffice
    int lut_shift;
    int e0;
    int sum_mode;
    int has_nulls;
    int has_scores;
};

// ---- device helpers ---------------------------------------------------------

#ifdef __CUDACC__

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

// REDUX: one instruction per warp-wide integer min / max on sm_80+
__device__ __forceinline__ u32 warp_min(u32 v) { return __reduce_min_sync(GRB_FULL, v); }
__device__ __forceinline__ u32 warp_max(u32 v) { return __reduce_max_sync(GRB_FULL, v); }
__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(GRB_FULL, v, o);
    return v;
}
__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(GRB_FULL, v, o);
    return v;
}

// f64 -> fixed point in units of 2^e0.  Exact by construction of e0 (every score is a multiple
// of 2^e0); zero, and anything the mode selection excluded, maps to 0.
__device__ __forceinline__ i128 to_fx(double x, int e0) {
    u64 b = (u64)__double_as_longlong(x);
    int ex = (int)((b >> 52) & 0x7ff);
    if (ex == 0 || ex == 0x7ff) return 0;
    u64 m = (b & ((1ull << 52) - 1)) | (1ull << 52);
    int sh = (ex - 1075) - e0;
    u128 v = sh >= 0 ? ((u128)m << sh) : ((u128)m >> (-sh));
    return (b >> 63) ? -(i128)v : (i128)v;
}

// fixed point -> nearest f64 (round half to even): the correctly rounded exact sum.  Values beyond 63 bits are
// shifted down to 62 significant bits with a sticky bit, so the one rounding of the integer -> f64 conversion is
// the rounding of the exact value.
__device__ __forceinline__ double fx_to_double(i128 v, int e0) {
    if (v == 0) return 0.0;
    const bool neg = v < 0;
    const u128 m = neg ? (u128)(-v) : (u128)v;
    const u64 hi = (u64)(m >> 64), lo = (u64)m;
    const int nbits = hi ? 128 - __clzll((long long)hi) : 64 - __clzll((long long)lo);
    int sh = 0;
    u64 q = lo;
    if (nbits > 63) {
        sh = nbits - 62;
        q = (u64)(m >> sh);
        if ((m & (((u128)1 << sh) - 1)) != 0) q |= 1;
    }
    double r = (double)(i64)q;
    // scale by 2^(sh+e0) in two exact steps (each factor is a normal power of two)
    const int e = sh + e0;
    const int e1 = e / 2, e2 = e - e1;
    r *= __longlong_as_double((long long)(u64)(1023 + e1) << 52);
    r *= __longlong_as_double((long long)(u64)(1023 + e2) << 52);
    return neg ? -r : r;
}

// the 96-bit two-plane prefix (hi:lo) as a sign-extended i128
__device__ __forceinline__ i128 sext96(u32 hi, u64 lo) { return (i128)(((u128)(u64)(i64)(int)hi << 64) | (u128)lo); }

// IEEE outcome of a left-to-right f64 sum that met np +inf, nn -inf and nan NaN scores: NaN beats everything,
// +inf and -inf together give NaN, one kind of infinity wins over any finite part.
__device__ __forceinline__ double nf_rule(double finite_sum, u32 np, u32 nn, u32 nan) {
    if (nan || (np && nn)) return __longlong_as_double(0x7ff8000000000000ll);
    if (np) return __longlong_as_double(0x7ff0000000000000ll);
    if (nn) return __longlong_as_double((long long)0xfff0000000000000ull);
    return finite_sum;
}

// order-preserving integer image of an f64 (and back): -inf < ... < -0.0 < +0.0 < ... < +inf
__device__ __forceinline__ u64 f64_key(double x) {
    u64 b = (u64)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | (1ull << 63));
}
__device__ __forceinline__ double key_f64(u64 k) {
    u64 b = (k >> 63) ? (k & ~(1ull << 63)) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#endif  // __CUDACC__

// ---- internal entry points between translation units -----------------------

// scan.cu: exclusive scan, out[i] = sum(in[0..i)), i in [0, n]; out has n+1 entries when
// write_total, else n.  Single-pass decoupled look-back.
int grb_scan_u32_to_u64(grb_ctx *ctx, const u32 *in, u64 *out, size_t n, bool write_total);
int grb_scan_u32_to_u32(grb_ctx *ctx, const u32 *in, u32 *out, size_t n, bool write_total);

// index.cu: build the member-sweep tables of an index (rank_a, score_sorted, fb) if they are not there yet.
int grb_index_prepare_members(grb_ctx *ctx, grb_index *ix);
// index.cu: the stabbing-minimum / -maximum tables and the block sparse tables that answer MIN / MAX without a sweep.
int grb_index_prepare_minmax(grb_ctx *ctx, grb_index *ix);
// index.cu: prefix sums of the sorted starts / ends (the closed form of OVERLAP_BP), built on first use.
int grb_index_prepare_bp(grb_ctx *ctx, grb_index *ix);
// index.cu: exact prefixes of score * coordinate (the closed form of WEIGHTED_MEAN); sets ix->wm_unavailable when the
// products do not fit the 128-bit fixed point (or the index has no exact sums) -- the caller then sweeps the members.
int grb_index_prepare_wm(grb_ctx *ctx, grb_index *ix);

// bucket.cu: left ranges in any order -- coordinate buckets, results written back through `rank`
int grb_left_disorder(grb_ctx *ctx, const u32 *ls, u64 first, u64 total, int *per_1024);
int grb_bucket_lefts(grb_ctx *ctx, const u64 *left_offsets, const u64 *extent, int nseq, const u32 *ls, const u32 *le, u32 *ls_b, u32 *le_b,
                     u32 *pos, u32 *inversions_dev);
int grb_unbucket_results(grb_ctx *ctx, u64 first, u64 total, int k, const double *out_b, const u8 *val_b, const u32 *pos, double *out, u8 *val);

// sort.cu: stable LSD radix sort of (u64 key, u32 value) pairs; skips digits that do not vary and
// the whole sort when the keys are already non-decreasing.  On return *keys_out / *vals_out point
// at whichever of the two buffers holds the result.
int grb_sort_pairs(grb_ctx *ctx, u64 *keys, u32 *vals, u64 *keys_alt, u32 *vals_alt, size_t n, u64 **keys_out,
                   u32 **vals_out, bool *was_sorted);
// sort.cu: stable argsort of u32 keys (read only): keys_sorted[i] = keys[perm[i]], equal keys in position order.
// tmp_keys / tmp_vals: scratch of n entries each.
int grb_argsort_u32(grb_ctx *ctx, const u32 *keys, u32 *keys_sorted, u32 *perm, u32 *tmp_keys, u32 *tmp_vals, size_t n);
