// reduce.cu -- composable reducers and Collapse.
//
// The reference lets a user fold any closure over a join (GRanges::map_joins, src/granges.rs:1176-1190).  The closures
// it ships -- FloatOperation::run (src/data/operations.rs:53-109), weighted_mean_score (examples/weighted_mean.rs:4-31),
// the covered-basepair sum of feature-density (src/commands.rs:839-843) -- are folds of value * weight over the
// overlapping rights.  grb_map_joins_reduce_batch takes such a fold described by its parts and maps every supported
// combination onto a reduction the device answers without visiting the members.  Collapse is text: the pairs are
// materialised (grb_left_overlaps: sort_ranges order), their scores gathered on the device, and the rows rendered on
// the host exactly like Rust's f64::to_string.
#include <string>
#include <vector>

#include "../host/format.hpp"
#include "common.cuh"

extern "C" int grb_map_joins_reduce_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                                          const uint32_t *ls, const uint32_t *le, const grb_reducer *reducers, int k, double *out,
                                          uint8_t *out_valid) {
    if (!ctx) return GRB_ERR_INVALID_ARG;
    if (!reducers || k < 1 || k > GRB_MAX_OPS) return grb_fail(ctx, GRB_ERR_INVALID_ARG, "map_joins_reduce: k must be in [1, %d]", GRB_MAX_OPS);
    int ops[GRB_MAX_OPS];
    for (int c = 0; c < k; c++) {
        const grb_reducer &r = reducers[c];
        int op = -1;
        const bool score = r.value == GRB_VALUE_SCORE, one = r.value == GRB_VALUE_ONE;
        const bool w1 = r.weight == GRB_WEIGHT_ONE, ww = r.weight == GRB_WEIGHT_OVERLAP_WIDTH;
        const bool zero = r.empty == GRB_EMPTY_ZERO, nov = r.empty == GRB_EMPTY_NOVALUE;
        if (r.fold == GRB_FOLD_SUM && r.divide == GRB_DIVIDE_NONE) {
            if (score && w1) op = zero ? GRB_OP_SUM : nov ? GRB_OP_SUM_NOT_EMPTY : -1;
            if (score && ww && zero) op = GRB_OP_WEIGHTED_SUM;
            if (one && w1 && zero) op = GRB_OP_COUNT;
            if (one && ww && zero) op = GRB_OP_OVERLAP_BP;
        } else if (r.fold == GRB_FOLD_SUM && r.divide == GRB_DIVIDE_BY_WEIGHTS) {
            if (score && w1 && nov) op = GRB_OP_MEAN;
            if (score && ww && zero) op = GRB_OP_WEIGHTED_MEAN;
        } else if ((r.fold == GRB_FOLD_MIN || r.fold == GRB_FOLD_MAX) && r.divide == GRB_DIVIDE_NONE && score && w1 && nov) {
            op = r.fold == GRB_FOLD_MIN ? GRB_OP_MIN : GRB_OP_MAX;
        }
        if (op < 0)
            return grb_fail(ctx, GRB_ERR_UNSUPPORTED,
                            "map_joins_reduce: column %d {value %d, weight %d, fold %d, divide %d, empty %d} has no member-free form "
                            "(materialise the pairs with grb_left_overlaps and fold them on the host)",
                            c, r.value, r.weight, r.fold, r.divide, r.empty);
        ops[c] = op;
    }
    return grb_map_joins_f64_batch(ctx, idx, left_offsets, nseq, ls, le, ops, k, out, out_valid);
}

// ---- Collapse --------------------------------------------------------------------------------------------------

struct grb_text {
    std::vector<uint64_t> offsets;
    std::string bytes;
};

namespace {

// score and validity of every pair (pos = the right's position in start order)
__global__ void k_pair_scores(const u32 *__restrict__ pos, u64 np, const double *__restrict__ score_a, const u32 *__restrict__ vbits_a,
                              double *__restrict__ val, u8 *__restrict__ ok) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const u32 j = pos[i];
    val[i] = score_a[j];
    ok[i] = vbits_a ? (u8)((vbits_a[j >> 5] >> (j & 31)) & 1u) : (u8)1;
}

}  // namespace

extern "C" int grb_collapse_f64(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl, grb_text **out) {
    if (!ctx || !out) return GRB_ERR_INVALID_ARG;
    *out = nullptr;
    if (idx && !idx->has_scores)
        return grb_fail(ctx, GRB_ERR_NO_DATA, "collapse: the right index has no data container (scores)");
    grb_text *t = new grb_text();
    t->offsets.assign(nl + 1, 0);
    if (nl == 0 || !idx || idx->n == 0) {
        *out = t;
        return GRB_OK;
    }
    std::vector<uint64_t> off(nl + 1);
    grb_pairs *p = nullptr;
    int rc = grb_left_overlaps(ctx, idx, ls, le, nl, off.data(), &p);
    if (rc != GRB_OK) {
        delete t;
        return rc;
    }
    const u64 np = p->np;
    std::vector<double> val(np);
    std::vector<u8> ok(np);
    if (np) {
        TempBuf dv, dk;
        rc = dv.alloc(ctx, np * 8);
        if (rc == GRB_OK) rc = dk.alloc(ctx, np);
        if (rc == GRB_OK) {
            k_pair_scores<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(p->pos, np, idx->score_a, idx->has_nulls ? idx->vbits_a : nullptr,
                                                                                dv.as<double>(), dk.as<u8>());
            ctx->launches++;
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(val.data(), dv.p, np * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ok.data(), dk.p, np, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = grb_fail(ctx, GRB_ERR_CUDA, "collapse: %s", cudaGetErrorString(e));
        }
    }
    grb_pairs_destroy(p);
    if (rc != GRB_OK) {
        delete t;
        return rc;
    }
    char tmp[448];
    for (size_t i = 0; i < nl; i++) {
        t->offsets[i] = t->bytes.size();
        bool first = true;
        for (u64 q = off[i]; q < off[i + 1]; q++) {
            if (!ok[q]) continue;  // None scores are dropped before the operation sees the group (src/commands.rs:526-531)
            if (!first) t->bytes.push_back(',');
            first = false;
            char *e = grh::format_f64(tmp, val[q]);
            t->bytes.append(tmp, (size_t)(e - tmp));
        }
    }
    t->offsets[nl] = t->bytes.size();
    *out = t;
    return GRB_OK;
}

extern "C" size_t grb_text_bytes(const grb_text *t) { return t ? t->bytes.size() : 0; }

extern "C" int grb_text_copy(const grb_text *t, uint64_t *offsets, char *bytes) {
    if (!t) return GRB_ERR_INVALID_ARG;
    if (offsets) memcpy(offsets, t->offsets.data(), t->offsets.size() * sizeof(uint64_t));
    if (bytes && !t->bytes.empty()) memcpy(bytes, t->bytes.data(), t->bytes.size());
    return GRB_OK;
}

extern "C" void grb_text_destroy(grb_text *t) { delete t; }
