// shard.cu -- host-side planner that splits one overlap join into independent units for N GPUs.
//
// Overlaps never cross seqnames (src/granges.rs:958-962 looks the right container up by the
// left's seqname), and within a seqname a run of lefts only needs the rights with
// r.start < max(l.end) and r.end > min(l.start).  So units are (seqname, contiguous left run,
// coordinate bounds of the rights it can touch); ranks own contiguous runs of units and never
// exchange data -- results are concatenated in (seqname, left rank) order (src/iterators.rs:43-76).
#include "common.cuh"

extern "C" int grb_shard_plan(const uint64_t *left_offsets, const uint64_t *right_counts, int nseq, const uint32_t *ls,
                              const uint32_t *le, int nranks, grb_shard *shards, int max_shards) {
    if (!left_offsets || nseq <= 0 || nranks <= 0 || !shards || max_shards <= 0) return GRB_ERR_INVALID_ARG;
    long double total = 0;
    for (int s = 0; s < nseq; s++) {
        u64 nl = left_offsets[s + 1] - left_offsets[s];
        if (nl == 0) continue;
        total += (long double)nl + (long double)(right_counts ? right_counts[s] : 0);
    }
    if (total == 0) return 0;
    const long double per_rank = total / nranks;
    int ns = 0;
    int rank = 0;
    long double used = 0;  // cost already given to `rank`
    for (int s = 0; s < nseq; s++) {
        const u64 lb = left_offsets[s], lend = left_offsets[s + 1];
        const u64 nl = lend - lb;
        if (nl == 0) continue;
        const long double cost_per_left = ((long double)nl + (long double)(right_counts ? right_counts[s] : 0)) / (long double)nl;
        u64 cur = lb;
        while (cur < lend) {
            long double room = per_rank - used;
            u64 take;
            if (rank == nranks - 1) {
                take = lend - cur;
            } else {
                long double fit = room / cost_per_left;
                take = fit >= (long double)(lend - cur) ? lend - cur : (u64)fit;
                // do not leave slivers: a piece under 1/64 of a rank's share stays with its neighbour
                if (lend - cur - take < (u64)(per_rank / cost_per_left / 64)) take = lend - cur;
                if (take == 0) {
                    rank++;
                    used = 0;
                    continue;
                }
            }
            if (ns >= max_shards) return GRB_ERR_INVALID_ARG;
            grb_shard *sh = &shards[ns++];
            sh->seq = s;
            sh->rank = rank;
            sh->left_begin = cur;
            sh->left_end = cur + take;
            u32 lo = 0xffffffffu, hi = 0;
            if (ls && le) {
                for (u64 i = cur; i < cur + take; i++) {
                    if (ls[i] < lo) lo = ls[i];
                    if (le[i] > hi) hi = le[i];
                }
            } else {
                lo = 0;
                hi = 0xffffffffu;
            }
            sh->right_lo = lo;
            sh->right_hi = hi;
            used += cost_per_left * (long double)take;
            cur += take;
            if (used >= per_rank * 0.999999L && rank < nranks - 1) {
                rank++;
                used = 0;
            }
        }
    }
    return ns;
}
