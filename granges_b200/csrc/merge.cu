// merge.cu -- the merging iterators (src/merging_iterators.rs:39-241, `granges merge` src/commands.rs:616-673,
// and the per-feature bookend merge of feature-density, src/commands.rs:818-826) on the device.
//
// The reference streams the records of a file: a record joins the open range when it is on the same
// seqname and last.distance_or_overlap(next) <= minimum_distance (src/traits.rs:89-108), extending
// last.end = max(last.end, next.end); otherwise the open range is emitted and the record opens a new one.
// A "run" below is a maximal stretch of consecutive records of one seqname (the caller splits the file).
//
//   * start-sorted runs, distance >= 0 (what `bedtools merge` accepts, and what every caller in the reference
//     feeds it): the open range's end is the running maximum of the ends, and that maximum is the maximum over
//     ALL earlier records of the run (an earlier group ended before this one began), so a record opens a group
//     iff start > prefix_max_end + distance.  One prefix-max scan (keyed by run, so it never crosses runs), one
//     flag scan, one emit pass.
//   * anything else (negative distance = minimum overlap, where a new group's end may be smaller than an
//     earlier end; unsorted input, where the rule also looks at the open range's start): one warp per run
//     walks the reference's state machine over coalesced 32-record loads.  Slow, exact, rarely needed.
//
// Reductions over the merged groups (MergingResultIterator's closure in `granges merge`, FloatOperation::run)
// reuse the overlap-join kernels: record i becomes the unit range [i, i+1) of a "position index" carrying its
// score, and group g the left range [first[g], first[g+1]) -- its overlapping rights are exactly its members.
#include <stdlib.h>

#include <vector>

#include "common.cuh"

struct grb_merged {
    grb_ctx *ctx;
    u64 n;       // input records
    u64 m;       // merged ranges
    int nruns;
    u32 *starts, *ends;  // device, m
    u32 *first;          // device, m + 1: input position of each group's first record; first[m] = n
    u64 *run_offsets_m;  // host, nruns + 1: merged ranges of run r are [run_offsets_m[r], run_offsets_m[r + 1])
};

namespace {

constexpr int MG_TPB = 256;
constexpr int MG_ITEMS = 4;
constexpr int MG_BLOCK = MG_TPB * MG_ITEMS;

__device__ __forceinline__ u32 find_run(const u64 *__restrict__ run_offsets, int nruns, u64 i) {
    int lo = 0, hi = nruns;  // largest r with run_offsets[r] <= i (empty runs are skipped by taking the last such r)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (run_offsets[mid] <= i)
            lo = mid;
        else
            hi = mid;
    }
    return (u32)lo;
}

struct MergeFlags {
    u32 unsorted;
    u32 invalid;
};

// key = run << 32 | end; also detects unsorted starts inside a run and invalid ranges
__global__ void k_merge_keys(const u32 *__restrict__ starts, const u32 *__restrict__ ends, const u64 *__restrict__ run_offsets,
                             int nruns, u64 n, u64 *__restrict__ keys, MergeFlags *fl) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 r = find_run(run_offsets, nruns, i);
    const u32 s = starts[i], e = ends[i];
    keys[i] = ((u64)r << 32) | e;
    if (e <= s && !fl->invalid) atomicOr(&fl->invalid, 1u);
    if (i > run_offsets[r] && starts[i - 1] > s && !fl->unsorted) atomicOr(&fl->unsorted, 1u);
}

__device__ __forceinline__ u64 warp_max64(u64 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const u64 t = __shfl_xor_sync(GRB_FULL, v, o);
        v = t > v ? t : v;
    }
    return v;
}

__global__ void __launch_bounds__(MG_TPB) k_block_max(const u64 *__restrict__ keys, u64 n, u64 *__restrict__ bmax) {
    __shared__ u64 s[MG_TPB / 32];
    const u64 base = (u64)blockIdx.x * MG_BLOCK;
    u64 m = 0;
#pragma unroll
    for (int k = 0; k < MG_ITEMS; k++) {
        const u64 i = base + (u64)k * MG_TPB + threadIdx.x;
        if (i < n) m = keys[i] > m ? keys[i] : m;
    }
    m = warp_max64(m);
    if (lane_id() == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < MG_TPB / 32; w++) t = s[w] > t ? s[w] : t;
        bmax[blockIdx.x] = t;
    }
}

// exclusive prefix maximum of nb values by one CTA
__global__ void __launch_bounds__(1024) k_prefix_max64_single(u64 *__restrict__ v, u64 nb) {
    __shared__ u64 s_m[1024];
    const u64 per = (nb + 1023) / 1024;
    const u64 lo = (u64)threadIdx.x * per;
    const u64 hi = lo + per < nb ? lo + per : nb;
    u64 m = 0;
    for (u64 i = lo; i < hi; i++) m = v[i] > m ? v[i] : m;
    s_m[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = 0;
        for (int k = 0; k < 1024; k++) {
            const u64 c = s_m[k];
            s_m[k] = run;
            run = c > run ? c : run;
        }
    }
    __syncthreads();
    u64 run = s_m[threadIdx.x];
    for (u64 i = lo; i < hi; i++) {
        const u64 c = v[i];
        v[i] = run;
        run = c > run ? c : run;
    }
}

// Per block: exclusive prefix max X[i] of the keys (carry = the blocks before), then
//   flag[i] = run(i) != run(X[i]) (first record of its run)  ||  start[i] > end(X[i]) + distance
//   incl_end[i] = the open range's end after record i = max(end[i], end(X[i]) if same run and not a new group)
// In the sorted, distance >= 0 case a new group's first end exceeds every earlier end of the run, so
// max(end[i], end(X[i])) is the open end whether or not i opens a group.
__global__ void __launch_bounds__(MG_TPB) k_merge_flags_sorted(const u32 *__restrict__ starts, const u64 *__restrict__ keys,
                                                               const u64 *__restrict__ bcarry, u64 n, long long distance,
                                                               u32 *__restrict__ flags, u32 *__restrict__ incl_end) {
    __shared__ u64 s_w[MG_TPB / 32];
    const u64 base = (u64)blockIdx.x * MG_BLOCK + (u64)threadIdx.x * MG_ITEMS;  // blocked: thread t owns 4 consecutive records
    u64 k[MG_ITEMS];
    u64 tmax = 0;
#pragma unroll
    for (int q = 0; q < MG_ITEMS; q++) {
        k[q] = base + q < n ? keys[base + q] : 0;
        tmax = k[q] > tmax ? k[q] : tmax;
    }
    // exclusive scan of the thread maxima across the CTA
    u64 incl = tmax;
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(GRB_FULL, incl, o);
        if (lane >= (u32)o) incl = t > incl ? t : incl;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u64 x = bcarry[blockIdx.x];
    for (u32 w = 0; w < warp; w++) x = s_w[w] > x ? s_w[w] : x;
    const u64 prev = __shfl_up_sync(GRB_FULL, incl, 1);
    if (lane) x = prev > x ? prev : x;
#pragma unroll
    for (int q = 0; q < MG_ITEMS; q++) {
        const u64 i = base + q;
        if (i < n) {
            const u32 run = (u32)(k[q] >> 32), e = (u32)k[q];
            const bool same_run = i > 0 && (u32)(x >> 32) == run && x != 0;
            bool fl = true;
            u32 open_end = e;
            if (same_run) {
                const u32 pe = (u32)x;
                fl = (long long)starts[i] > (long long)pe + distance;
                open_end = pe > e ? pe : e;
            }
            flags[i] = fl ? 1u : 0u;
            incl_end[i] = open_end;
        }
        x = k[q] > x ? k[q] : x;
    }
}

// distance_or_overlap -- src/traits.rs:89-108 (a = the open range, b = the next record)
__device__ __forceinline__ long long distance_or_overlap(u32 as, u32 ae, u32 bs, u32 be) {
    if (ae > bs && as < be) {
        const u32 os = as > bs ? as : bs, oe = ae < be ? ae : be;
        const u32 ov = os >= oe ? 0u : oe - os;  // overlap_width, src/traits.rs:73-80
        return -(long long)ov;
    }
    if (ae == bs || as == be) return 0;
    const u32 d1 = bs > ae ? bs - ae : 0u, d2 = as > be ? as - be : 0u;  // saturating_sub
    return (long long)(d1 > d2 ? d1 : d2);
}

// The reference's state machine, one warp per run: 32 records are loaded at once, then replayed in order.
__global__ void __launch_bounds__(32) k_merge_flags_sequential(const u32 *__restrict__ starts, const u32 *__restrict__ ends,
                                                               const u64 *__restrict__ run_offsets, int nruns, long long distance,
                                                               u32 *__restrict__ flags, u32 *__restrict__ incl_end) {
    const int r = blockIdx.x;
    if (r >= nruns) return;
    const u64 lo = run_offsets[r], hi = run_offsets[r + 1];
    const u32 lane = lane_id();
    u32 cur_s = 0, cur_e = 0;
    bool open = false;
    for (u64 base = lo; base < hi; base += 32) {
        const u64 i = base + lane;
        u32 s = 0, e = 0;
        if (i < hi) {
            s = starts[i];
            e = ends[i];
        }
        const u32 cnt = hi - base < 32 ? (u32)(hi - base) : 32u;
        u32 my_flag = 0, my_end = 0;
        for (u32 q = 0; q < cnt; q++) {
            const u32 qs = __shfl_sync(GRB_FULL, s, q), qe = __shfl_sync(GRB_FULL, e, q);
            bool fl;
            if (open && distance_or_overlap(cur_s, cur_e, qs, qe) <= distance) {
                cur_e = cur_e > qe ? cur_e : qe;
                fl = false;
            } else {
                cur_s = qs;
                cur_e = qe;
                open = true;
                fl = true;
            }
            if (q == lane) {
                my_flag = fl ? 1u : 0u;
                my_end = cur_e;
            }
        }
        if (i < hi) {
            flags[i] = my_flag;
            incl_end[i] = my_end;
        }
    }
}

__global__ void k_merge_emit(const u32 *__restrict__ starts, const u32 *__restrict__ flags, const u32 *__restrict__ incl_end,
                             const u64 *__restrict__ pos, u64 n, u64 m, u32 *__restrict__ mstarts, u32 *__restrict__ mends,
                             u32 *__restrict__ first) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) {
        const u64 g = pos[i];
        mstarts[g] = starts[i];
        first[g] = (u32)i;
        if (g) mends[g - 1] = incl_end[i - 1];
    }
    if (i == n - 1) {
        mends[m - 1] = incl_end[i];
        first[m] = (u32)n;
    }
}

__global__ void k_gather_u64(const u64 *__restrict__ src, const u64 *__restrict__ at, int k, u64 *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) dst[i] = src[at[i]];
}

__global__ void k_iota_ranges(u64 n, u32 *__restrict__ s, u32 *__restrict__ e) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        s[i] = (u32)i;
        e[i] = (u32)i + 1u;
    }
}

__global__ void k_shift_copy(const u32 *__restrict__ src, u64 n, u32 *__restrict__ dst) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i + 1];
}

inline unsigned blocks_for(u64 n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

}  // namespace

extern "C" {

int grb_merge_ranges(grb_ctx *ctx, const uint32_t *starts, const uint32_t *ends, const uint64_t *run_offsets, int nruns,
                     int64_t distance, grb_merged **out) {
    if (!ctx || !out) return GRB_ERR_INVALID_ARG;
    *out = nullptr;
    if (nruns < 0 || (nruns && !run_offsets)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merge: NULL run_offsets");
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    const u64 n = nruns ? run_offsets[nruns] : 0;
    for (int r = 0; r < nruns; r++)
        if (run_offsets[r + 1] < run_offsets[r]) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merge: run_offsets must be non-decreasing");
    if (nruns && run_offsets[0] != 0) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merge: run_offsets[0] must be 0");
    if (n && (!starts || !ends)) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merge: NULL range buffer");
    if (n >= 0x7fffffffull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "merge: more than 2^31-2 records per call");
    grb_merged *mg = (grb_merged *)calloc(1, sizeof(grb_merged));
    if (!mg) return grb_fail(ctx, GRB_ERR_OOM, "merge: host allocation failed");
    mg->ctx = ctx;
    mg->n = n;
    mg->nruns = nruns;
    mg->run_offsets_m = (u64 *)calloc((size_t)nruns + 1, sizeof(u64));
    if (!mg->run_offsets_m) {
        free(mg);
        return grb_fail(ctx, GRB_ERR_OOM, "merge: host allocation failed");
    }
    if (n == 0) {
        *out = mg;
        return GRB_OK;
    }
    int rc = GRB_OK;
    do {
        InBuf d_s, d_e;
        if ((rc = d_s.stage(ctx, starts, n * 4))) break;
        if ((rc = d_e.stage(ctx, ends, n * 4))) break;
        // run offsets always come from the host
        TempBuf ro;
        if ((rc = ro.alloc(ctx, ((size_t)nruns + 1) * 8))) break;
        if (cudaMemcpyAsync(ro.p, run_offsets, ((size_t)nruns + 1) * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "merge: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        TempBuf keys, flags, iend, pos, fl;
        if ((rc = keys.alloc(ctx, n * 8))) break;
        if ((rc = flags.alloc(ctx, n * 4))) break;
        if ((rc = iend.alloc(ctx, n * 4))) break;
        if ((rc = pos.alloc(ctx, (n + 1) * 8))) break;
        if ((rc = fl.alloc(ctx, sizeof(MergeFlags)))) break;
        if (cudaMemsetAsync(fl.p, 0, sizeof(MergeFlags), ctx->stream) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "merge: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        k_merge_keys<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_s.as<u32>(), d_e.as<u32>(), ro.as<u64>(), nruns, n, keys.as<u64>(),
                                                                 fl.as<MergeFlags>());
        GRB_LAUNCHED_BREAK(ctx, rc)
        MergeFlags hf;
        if (cudaMemcpyAsync(&hf, fl.p, sizeof(hf), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "merge: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (hf.invalid) {
            rc = grb_fail(ctx, GRB_ERR_INVALID_RANGE, "merge: a range has end <= start");
            break;
        }
        const char *force = getenv("GRB_MERGE_MODE");  // debug: 1 = always the sequential state machine
        if (distance >= 0 && !hf.unsorted && !(force && atoi(force) == 1)) {
            const u64 nb = (n + MG_BLOCK - 1) / MG_BLOCK;
            TempBuf bmax;
            if ((rc = bmax.alloc(ctx, nb * 8))) break;
            k_block_max<<<(unsigned)nb, MG_TPB, 0, ctx->stream>>>(keys.as<u64>(), n, bmax.as<u64>());
            GRB_LAUNCHED_BREAK(ctx, rc)
            k_prefix_max64_single<<<1, 1024, 0, ctx->stream>>>(bmax.as<u64>(), nb);
            GRB_LAUNCHED_BREAK(ctx, rc)
            k_merge_flags_sorted<<<(unsigned)nb, MG_TPB, 0, ctx->stream>>>(d_s.as<u32>(), keys.as<u64>(), bmax.as<u64>(), n, (long long)distance,
                                                                          flags.as<u32>(), iend.as<u32>());
            GRB_LAUNCHED_BREAK(ctx, rc)
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {  // bmax goes out of scope
                rc = grb_fail(ctx, GRB_ERR_CUDA, "merge: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
        } else {
            k_merge_flags_sequential<<<(unsigned)nruns, 32, 0, ctx->stream>>>(d_s.as<u32>(), d_e.as<u32>(), ro.as<u64>(), nruns,
                                                                             (long long)distance, flags.as<u32>(), iend.as<u32>());
            GRB_LAUNCHED_BREAK(ctx, rc)
        }
        if ((rc = grb_scan_u32_to_u64(ctx, flags.as<u32>(), pos.as<u64>(), n, true))) break;
        // merged ranges in all, and per run
        TempBuf rom;
        if ((rc = rom.alloc(ctx, ((size_t)nruns + 1) * 8))) break;
        k_gather_u64<<<blocks_for((u64)nruns + 1, 128), 128, 0, ctx->stream>>>(pos.as<u64>(), ro.as<u64>(), nruns + 1, rom.as<u64>());
        GRB_LAUNCHED_BREAK(ctx, rc)
        if (cudaMemcpyAsync(mg->run_offsets_m, rom.p, ((size_t)nruns + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "merge: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        const u64 m = mg->run_offsets_m[nruns];
        mg->m = m;
        if (grb_dev_malloc(ctx, (void **)&mg->starts, m * 4) != cudaSuccess || grb_dev_malloc(ctx, (void **)&mg->ends, m * 4) != cudaSuccess ||
            grb_dev_malloc(ctx, (void **)&mg->first, (m + 1) * 4 + 16) != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_OOM, "merge: device allocation failed");
            break;
        }
        k_merge_emit<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_s.as<u32>(), flags.as<u32>(), iend.as<u32>(), pos.as<u64>(), n, m,
                                                                 mg->starts, mg->ends, mg->first);
        GRB_LAUNCHED_BREAK(ctx, rc)
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            rc = grb_fail(ctx, GRB_ERR_CUDA, "merge: %s", cudaGetErrorString(e));
            break;
        }
    } while (0);
    if (rc != GRB_OK) {
        grb_merged_destroy(mg);
        return rc;
    }
    *out = mg;
    return GRB_OK;
}

size_t grb_merged_len(const grb_merged *m) { return m ? (size_t)m->m : 0; }

int grb_merged_copy(const grb_merged *mg, uint32_t *starts, uint32_t *ends, uint32_t *first, uint64_t *run_offsets_merged) {
    if (!mg) return GRB_ERR_INVALID_ARG;
    grb_ctx *ctx = mg->ctx;
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mg->m) {
        if (starts) GRB_CUDA(ctx, cudaMemcpyAsync(starts, mg->starts, mg->m * 4, cudaMemcpyDefault, ctx->stream));
        if (ends) GRB_CUDA(ctx, cudaMemcpyAsync(ends, mg->ends, mg->m * 4, cudaMemcpyDefault, ctx->stream));
        if (first) GRB_CUDA(ctx, cudaMemcpyAsync(first, mg->first, (mg->m + 1) * 4, cudaMemcpyDefault, ctx->stream));
    } else if (first) {
        first[0] = 0;
    }
    if (run_offsets_merged)
        for (int r = 0; r <= mg->nruns; r++) run_offsets_merged[r] = mg->run_offsets_m[r];
    GRB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

const uint32_t *grb_merged_dev_starts(const grb_merged *m) { return m ? m->starts : nullptr; }
const uint32_t *grb_merged_dev_ends(const grb_merged *m) { return m ? m->ends : nullptr; }

int grb_merged_map_f64(grb_ctx *ctx, const grb_merged *mg, const double *scores, const uint8_t *valid, const int *ops, int k,
                       double *out, uint8_t *out_valid) {
    if (!ctx || !mg) return GRB_ERR_INVALID_ARG;
    if (mg->ctx != ctx) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merged_map: handle belongs to another context");
    if (mg->m == 0) return GRB_OK;
    if (!scores || !ops || !out || !out_valid) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merged_map: NULL buffer");
    for (int c = 0; c < k; c++)
        if (ops[c] == GRB_OP_OVERLAP_BP || ops[c] == GRB_OP_WEIGHTED_MEAN)
            return grb_fail(ctx, GRB_ERR_INVALID_ARG, "merged_map: operation %d is a join reduction, not a FloatOperation", ops[c]);
    GRB_CUDA(ctx, cudaSetDevice(ctx->device));
    const u64 n = mg->n, m = mg->m;
    grb_index *pix = nullptr;
    int rc;
    {
        TempBuf ps, pe;
        GRB_TRY(ps.alloc(ctx, n * 4));
        GRB_TRY(pe.alloc(ctx, n * 4));
        k_iota_ranges<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(n, ps.as<u32>(), pe.as<u32>());
        GRB_LAUNCHED(ctx);
        rc = grb_index_build(ctx, ps.as<u32>(), pe.as<u32>(), nullptr, n, &pix);
    }
    if (rc == GRB_OK) rc = grb_index_set_scores(pix, scores, valid, n);
    if (rc == GRB_OK) {
        TempBuf le;
        rc = le.alloc(ctx, m * 4);
        if (rc == GRB_OK) {
            k_shift_copy<<<blocks_for(m, 256), 256, 0, ctx->stream>>>(mg->first, m, le.as<u32>());
            ctx->launches++;
            if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) rc = grb_fail(ctx, GRB_ERR_CUDA, "merged_map: kernel launch: %s", cudaGetErrorString(e));
            const uint64_t off[2] = {0, m};
            const grb_index *cp = pix;
            if (rc == GRB_OK) rc = grb_map_joins_f64_batch(ctx, &cp, off, 1, mg->first, le.as<u32>(), ops, k, out, out_valid);
            if (rc == GRB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                rc = grb_fail(ctx, GRB_ERR_CUDA, "merged_map: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    grb_index_destroy(pix);
    return rc;
}

void grb_merged_destroy(grb_merged *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    grb_dev_free(m->ctx, m->starts);
    grb_dev_free(m->ctx, m->ends);
    grb_dev_free(m->ctx, m->first);
    free(m->run_offsets_m);
    free(m);
}

}  // extern "C"
