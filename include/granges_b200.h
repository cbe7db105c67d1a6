/*
 * granges_b200.h -- C ABI of the B200-native overlap-join hot path.
 *
 * This is the drop-in boundary: the entry points below are what a thin
 * `extern "C"` crate inside vsbuffalo/granges (v0.4.0) would bind from
 * src/ranges/coitrees.rs, src/granges.rs, src/join.rs and src/commands.rs
 * (see INTEGRATION.md for the Rust-side stub).  Every function cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - All functions return an int status: 0 = GRB_OK, negative = error
 *     (grb_status).  The message of the last error on a context is returned
 *     by grb_last_error(ctx); the error codes mirror GRangesError variants
 *     (src/error.rs:72-176) or the reference's panics where it panics.
 *   - Plain pointers and sizes only.  The caller owns every buffer it passes
 *     in; the library owns device memory behind opaque handles that are freed
 *     by the matching *_destroy.
 *   - Range buffers are per seqname (one grb_index per seqname mirrors
 *     GenomeMap<COITrees<M>>, src/granges.rs:71-78); the *_batch entry points
 *     take the seqnames of one GRanges concatenated in GenomeMap order
 *     (src/iterators.rs:43-76) so one launch covers the whole genome.
 *   - ls/le/out/... pointers may be HOST pointers (pageable or pinned) or
 *     DEVICE pointers of the context's device; the library detects which
 *     (cudaPointerGetAttributes) and stages host buffers itself: large
 *     pageable buffers (a Rust Vec, a numpy array) travel through the
 *     library's own ring of page-locked staging buffers.  With host
 *     buffers a call returns after its results are in the caller's buffers;
 *     with device buffers it returns after enqueueing on the context's stream.
 *   - A grb_ctx is bound to one CUDA device and one stream and is not
 *     thread-safe (one context per GPU per thread).
 *   - Lifetimes: handles hold plain back-pointers.  A context must outlive
 *     every handle created on it (indexes, pairs, ranges, merged sets:
 *     destroy them first), and an index must outlive the grb_pairs made from
 *     it (grb_pairs_copy / grb_pairs_overlap_widths read the index).
 *   - Invalid left ranges (end <= start, end > 2^31-1: the reference cannot
 *     even construct them, src/ranges/mod.rs:30) are answered as empty groups
 *     and flagged on the device.  Calls with HOST buffers check the flag
 *     before they return (GRB_ERR_INVALID_RANGE); calls with DEVICE buffers
 *     only enqueue work, so the error surfaces at the next call that
 *     synchronises -- grb_ctx_sync, or any call with host buffers.
 *   - Work is ordered on the context's stream only: device buffers produced
 *     on another stream need that stream synchronised (or the context bound
 *     to it with grb_ctx_set_stream) before they are passed in.
 *   - Coordinates are 0-based half-open u32 with end > start and
 *     end <= 2^31-1 (the reference converts to i32 and panics beyond that,
 *     src/ranges/coitrees.rs:53-54,173-178).
 *   - There is no CPU fallback: without a usable CUDA device every call that
 *     needs one fails with GRB_ERR_CUDA.
 */
#ifndef GRANGES_B200_H
#define GRANGES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRB_ABI_VERSION 1

typedef enum grb_status {
    GRB_OK = 0,
    GRB_ERR_INVALID_ARG = -1,   /* NULL where a buffer is required, bad op code, k out of range */
    GRB_ERR_INVALID_RANGE = -2, /* end <= start (RangeIndexed::new asserts, src/ranges/mod.rs:30,109) or
                                   end > i32::MAX (src/ranges/coitrees.rs:53-54 `expect`) */
    GRB_ERR_NO_DATA = -3,       /* GRangesError::NoDataContainer (src/granges.rs:955): reduction needs scores */
    GRB_ERR_CUDA = -4,          /* CUDA runtime failure; message holds cudaGetErrorString */
    GRB_ERR_OOM = -5,           /* device or host allocation failed */
    GRB_ERR_NAN_SCORE = -6,     /* median over NaN: the reference panics in partial_cmp().unwrap()
                                   (src/data/operations.rs:22-29) */
    GRB_ERR_INDEX_RANGE = -7,   /* data index >= n_data (reference: slice index panic, src/data/vec.rs:20-22) */
    GRB_ERR_UNSUPPORTED = -8,   /* e.g. more than 2^32-2 ranges on one seqname */
} grb_status;

/* FloatOperation (src/data/operations.rs:35-51) plus the two join-metadata
 * reductions closures commonly use (src/join.rs:153-163).  Collapse yields text, not a
 * number: see grb_collapse_f64 below. */
typedef enum grb_op {
    GRB_OP_SUM = 0,           /* empty -> 0.0            operations.rs:60-63 */
    GRB_OP_SUM_NOT_EMPTY = 1, /* empty -> NoValue        operations.rs:64-70 */
    GRB_OP_MIN = 2,           /* finite values only      operations.rs:71-78 */
    GRB_OP_MAX = 3,           /* finite values only      operations.rs:79-86 */
    GRB_OP_MEAN = 4,          /* empty -> NoValue        operations.rs:87-95 */
    GRB_OP_MEDIAN = 5,        /* empty -> NoValue        operations.rs:17-32,96-98 */
    GRB_OP_COUNT = 6,         /* LeftGroupedJoin::num_overlaps, join.rs:153 (counts None-score rights too) */
    GRB_OP_OVERLAP_BP = 7,    /* sum of overlap_widths, join.rs:158-163, commands.rs:839-843 */
    GRB_OP_WEIGHTED_MEAN = 8, /* sum(score * overlap_width) / sum(overlap_width) over the non-missing scores, 0.0 if none:
                                 the closure of examples/weighted_mean.rs:4-31 (join.rs:158-163 overlap_widths) */
    GRB_OP_WEIGHTED_SUM = 9,  /* sum(score * overlap_width) over the non-missing scores, 0.0 if none (the numerator above) */
    GRB_OP__COUNT = 10
} grb_op;

#define GRB_MAX_OPS 16

typedef struct grb_ctx grb_ctx;
typedef struct grb_index grb_index;
typedef struct grb_pairs grb_pairs;
typedef struct grb_ranges grb_ranges;
typedef struct grb_merged grb_merged;

/* ---- context ----------------------------------------------------------- */

int grb_abi_version(void);
/* Bind a context to CUDA device `device` (creates its own non-blocking stream). */
int grb_ctx_create(int device, grb_ctx **out);
/* Run on a caller-provided cudaStream_t instead (a null handle is CUDA's default stream). */
int grb_ctx_set_stream(grb_ctx *ctx, void *cuda_stream);
/* Back to the stream the context created for itself. */
int grb_ctx_use_own_stream(grb_ctx *ctx);
/* Left ranges may come in any order (GRanges::left_overlaps walks the left container as it is, src/granges.rs:935-984)
 * and every order gives the same results.  Lefts whose neighbours are close in coordinate (sorted BED files,
 * from_windows) run fastest; lefts found to be out of order are dropped into coordinate buckets on the device first
 * and their results written back to the rows they came from (about 3x the sorted time instead of 20x).
 *   AUTO      host buffers: a few thousand adjacent pairs are sampled on the host.  Device buffers: no extra work is
 *             put on the stream -- the kernels report how many tiles fell off the staged path, and a call that repeats
 *             the previous call's left buffer uses that report (the first call on shuffled device lefts is slow).
 *   SORTED    never bucket.        UNSORTED  always bucket. */
typedef enum grb_left_order { GRB_LEFT_ORDER_AUTO = 0, GRB_LEFT_ORDER_SORTED = 1, GRB_LEFT_ORDER_UNSORTED = 2 } grb_left_order;
int grb_ctx_set_left_order(grb_ctx *ctx, int order);
/* Block until everything enqueued on the context's stream has finished. */
int grb_ctx_sync(grb_ctx *ctx);
void grb_ctx_destroy(grb_ctx *ctx);
const char *grb_last_error(const grb_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t grb_ctx_launch_count(const grb_ctx *ctx);
/* Device properties the host-side shard planner needs. */
int grb_ctx_sm_count(const grb_ctx *ctx);

/* ---- right-hand index: replaces COITrees<M> ---------------------------- */

/* COITrees::from(VecRanges<R>) -- src/ranges/coitrees.rs:78-84 via GRanges::into_coitrees
 * (src/granges.rs:1365-1402).  Builds ONE seqname's index: rights sorted by (start, end,
 * data index) as a structure of arrays, the ends additionally sorted on their own, and a
 * block-max-end array; this replaces the interval tree.  data_idx NULL means 0..n-1 (insertion
 * order, what push_range assigns, src/granges.rs:733-754; also used for COITreesEmpty).
 * n may be 0.  Fails with GRB_ERR_INVALID_RANGE like the reference's assert / expect. */
int grb_index_build(grb_ctx *ctx, const uint32_t *starts, const uint32_t *ends, const uint64_t *data_idx, size_t n,
                    grb_index **out);
void grb_index_destroy(grb_index *idx);
/* COITrees::len -- src/ranges/coitrees.rs:67-69. */
size_t grb_index_len(const grb_index *idx);
/* The right data container reduced to Option<f64> per data index: what the map_data closure of
 * granges_map produces (src/commands.rs:514-517).  valid NULL = every score is Some; otherwise
 * valid[i] == 0 marks None (BED "." scores, src/io/parsers/bed/mod.rs:37-73).  Also builds the
 * exact fixed-point prefix sums SUM / MEAN are answered from. */
int grb_index_set_scores(grb_index *idx, const double *by_data_idx, const uint8_t *valid, size_t n_data);
/* GRanges::into_coitrees for the whole right-hand GRanges (src/granges.rs:1365-1402 loops over the
 * seqnames) plus the score extraction: builds nseq indexes (data index = position within the seqname's
 * arrays), copying seqname s+1 to the device while seqname s is being indexed.  scores / valid may be NULL
 * (no data container), as may individual entries.  n[s] may be 0.  out receives nseq handles. */
int grb_index_build_batch(grb_ctx *ctx, int nseq, const uint32_t *const *starts, const uint32_t *const *ends,
                          const size_t *n, const double *const *scores, const uint8_t *const *valid, grb_index **out);
/* Sorted iteration, COITrees::iter_ranges (src/ranges/coitrees.rs:154-170): copies the index
 * order out (any pointer may be NULL). */
int grb_index_copy_sorted(const grb_index *idx, uint32_t *starts, uint32_t *ends, uint64_t *data_idx);
/* 1 if SUM/MEAN on this index are served by the exact prefix-sum path, 0 if by the sweep. */
int grb_index_exact_sums(const grb_index *idx);

/* ---- queries ----------------------------------------------------------- */

/* COITrees::count_overlaps -- src/ranges/coitrees.rs:60-64, for nl left ranges at once.
 * idx NULL (right has no such seqname) gives zeros. */
int grb_count_overlaps(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                       uint32_t *counts);

/* filter_overlaps / antifilter_overlaps -- src/granges.rs:1416-1441,1485-1508,1524-1609:
 * keep[i] = (count_overlaps > 0) XOR anti; idx NULL: the semi-join drops, the anti-join keeps
 * (src/granges.rs:1433-1437,1580-1593).  n_kept (may be NULL) receives the number kept. */
int grb_filter_overlaps(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                        int anti, uint8_t *keep, uint64_t *n_kept);

/* LeftOverlaps::left_overlaps -- src/traits.rs:44-48, impls src/granges.rs:839-1034: the left
 * grouped join in CSR form.  offsets[nl+1] (u64): group i is pairs [offsets[i], offsets[i+1]);
 * every left gets a group (left join), also with idx NULL.  Within a group pairs are in
 * LeftGroupedJoin::sort_ranges order (start, end, index) -- src/join.rs:121-128. */
int grb_left_overlaps(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                      uint64_t *offsets, grb_pairs **out);
size_t grb_pairs_len(const grb_pairs *p);
/* RangeTuple fields of every pair (src/join.rs:17): right data index, right start, right end.
 * Any pointer may be NULL. */
int grb_pairs_copy(const grb_pairs *p, uint64_t *right_data_idx, uint32_t *r_start, uint32_t *r_end);
/* LeftGroupedJoin::overlap_widths (src/join.rs:158-163): overlap bp of every pair. */
int grb_pairs_overlap_widths(const grb_pairs *p, uint32_t *widths);
void grb_pairs_destroy(grb_pairs *p);

/* left_overlaps followed by map_joins with built-in reductions, fused -- no pair is
 * materialised.  Replaces src/granges.rs:935-984 + :1176-1190 + src/join.rs:343-361 +
 * the closure of granges_map (src/commands.rs:524-542) + FloatOperation::run
 * (src/data/operations.rs:53-109).  out is nl x k row-major f64, out_valid nl x k bytes:
 * 1 = DatumType::Float64(out), 0 = DatumType::NoValue.  idx NULL: every group is empty. */
int grb_map_joins_f64(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl,
                      const int *ops, int k, double *out, uint8_t *out_valid);

/* Whole-GRanges variants: seqname s (GenomeMap order) owns lefts [left_offsets[s], left_offsets[s+1])
 * of ls/le and rows of the outputs; idx[s] may be NULL.  One launch sequence for all seqnames.
 * left_offsets is a host array of nseq+1 entries. */
int grb_count_overlaps_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                             const uint32_t *ls, const uint32_t *le, uint32_t *counts);
int grb_filter_overlaps_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                              const uint32_t *ls, const uint32_t *le, int anti, uint8_t *keep, uint64_t *n_kept);
/* The same semi- / anti-join as an order-preserving stream compaction (keep flags -> scan -> scatter on the
 * device): kept_index receives the ascending global indices (into ls / le) of the lefts that are kept --
 * the rows of the filtered GRanges, data included, are the input rows at those indices
 * (src/granges.rs:1485-1508, 1547-1609).  kept_index must hold up to left_offsets[nseq] entries. */
int grb_filter_overlaps_select(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                               const uint32_t *ls, const uint32_t *le, int anti, uint64_t *kept_index, uint64_t *n_kept);
int grb_map_joins_f64_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                            const uint32_t *ls, const uint32_t *le, const int *ops, int k, double *out,
                            uint8_t *out_valid);

/* Composable reducers: the closures users write over a join -- examples/weighted_mean.rs:4-31,
 * src/commands.rs:839-843 -- are folds of  value(right) * weight(left, right)  over the overlapping rights.  Instead of
 * one enum value per closure a reducer is described by its parts, and every column of one call may have its own:
 *   value   SCORE: the right's score, rights whose score is None are skipped (src/commands.rs:526-531)
 *           ONE:   1 per overlapping right, scored or not
 *   weight  ONE, or OVERLAP_WIDTH = LeftGroupedJoin::overlap_widths (src/join.rs:158-163)
 *   fold    SUM, MIN, MAX (MIN / MAX see finite values only, like FloatOperation::Min / Max; they take weight ONE)
 *   divide  NONE, or BY_WEIGHTS: divide the sum by the summed weights of the rights that contributed (a mean)
 *   empty   what a left without contributing rights gets: ZERO (0.0, a value) or NOVALUE (DatumType::NoValue)
 * e.g. {SCORE, OVERLAP_WIDTH, SUM, BY_WEIGHTS, ZERO} is the weighted mean of examples/weighted_mean.rs,
 * {ONE, OVERLAP_WIDTH, SUM, NONE, ZERO} the covered basepairs of feature-density, {SCORE, ONE, SUM, BY_WEIGHTS, NOVALUE}
 * FloatOperation::Mean.  Every supported combination is answered without visiting the members (prefix sums, closed forms,
 * stabbing tables); a combination with no such form returns GRB_ERR_UNSUPPORTED. */
typedef enum grb_red_value { GRB_VALUE_SCORE = 0, GRB_VALUE_ONE = 1 } grb_red_value;
typedef enum grb_red_weight { GRB_WEIGHT_ONE = 0, GRB_WEIGHT_OVERLAP_WIDTH = 1 } grb_red_weight;
typedef enum grb_red_fold { GRB_FOLD_SUM = 0, GRB_FOLD_MIN = 1, GRB_FOLD_MAX = 2 } grb_red_fold;
typedef enum grb_red_divide { GRB_DIVIDE_NONE = 0, GRB_DIVIDE_BY_WEIGHTS = 1 } grb_red_divide;
typedef enum grb_red_empty { GRB_EMPTY_ZERO = 0, GRB_EMPTY_NOVALUE = 1 } grb_red_empty;
typedef struct grb_reducer {
    int value, weight, fold, divide, empty;
} grb_reducer;
int grb_map_joins_reduce_batch(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                               const uint32_t *ls, const uint32_t *le, const grb_reducer *reducers, int k, double *out,
                               uint8_t *out_valid);

/* FloatOperation::Collapse (src/data/operations.rs:99-106): the non-missing scores of every group as text, joined by
 * ",", each printed like Rust's f64::to_string (shortest round-trip digits, positional).  The reference concatenates in
 * the tree's visit order, which it leaves unspecified (src/join.rs:97); here the order is DEFINED as
 * LeftGroupedJoin::sort_ranges order -- (start, end, data index), src/join.rs:121-128 -- so the output is reproducible.
 * A left without scored overlaps gets the empty string (DatumType::String("")).  Row i is bytes
 * [offsets[i], offsets[i+1]) of the text; one seqname per call (idx NULL: every row empty). */
typedef struct grb_text grb_text;
int grb_collapse_f64(grb_ctx *ctx, const grb_index *idx, const uint32_t *ls, const uint32_t *le, size_t nl, grb_text **out);
size_t grb_text_bytes(const grb_text *t);
/* offsets: nl + 1 entries; bytes: grb_text_bytes(t) characters, no terminator.  Host buffers. */
int grb_text_copy(const grb_text *t, uint64_t *offsets, char *bytes);
void grb_text_destroy(grb_text *t);

/* The same call with the validity as BITS: bit (i * k + c) of out_valid_bits (word (i * k + c) / 32, u32 words,
 * ceil(nl * k / 32) of them) is 1 for DatumType::Float64, 0 for DatumType::NoValue -- the NL * K / 8 bytes SURVEY 8(d)
 * budgets for the valid flags instead of NL * K.  A single prefix-sum reduction (sum / sum-not-empty / mean / count)
 * writes the bits from the kernel itself; other calls pack them on the device.  left_offsets[0] must be 0.  Bits past
 * nl * k in the last word are unspecified. */
int grb_map_joins_f64_batch_bits(grb_ctx *ctx, const grb_index *const *idx, const uint64_t *left_offsets, int nseq,
                                 const uint32_t *ls, const uint32_t *le, const int *ops, int k, double *out,
                                 uint32_t *out_valid_bits);

/* `granges map` in one call for host-resident inputs -- src/commands.rs:477-547 (into_coitrees -> map_data ->
 * left_overlaps -> map_joins): grb_index_build_batch + grb_map_joins_f64_batch, pipelined per seqname so that the
 * results of the first seqnames travel back while the rights of the later ones are still being uploaded and indexed.
 * Same results as the two calls; the indexes are internal and gone on return.  r_scores / r_valid as in
 * grb_index_build_batch; r_n[s] == 0 means the right file never mentions the seqname. */
int grb_map_joins_whole_f64(grb_ctx *ctx, int nseq, const uint32_t *const *r_starts, const uint32_t *const *r_ends,
                            const size_t *r_n, const double *const *r_scores, const uint8_t *const *r_valid,
                            const uint64_t *left_offsets, const uint32_t *ls, const uint32_t *le, const int *ops, int k,
                            double *out, uint8_t *out_valid);

/* The same call on several GPUs of one box, in one process.  Overlaps never cross seqnames (src/granges.rs:958-962), so
 * the seqnames are cut into `ndev` contiguous runs of near-equal transfer volume; one host thread and one context per GPU
 * (created on first use, kept for later calls) runs grb_map_joins_whole_f64 on its run, and every run writes its own
 * rows of the caller's out / out_valid -- the host-side concatenation in the reference's row order (seqnames in GenomeMap
 * order, then left insertion order: src/iterators.rs:43-76).  No data moves between the GPUs.  Bit-identical to the
 * single-GPU call.  Host buffers only (pageable or page-locked).  The message of a failure is copied to err (may be NULL). */
int grb_map_joins_whole_f64_mgpu(const int *devices, int ndev, int nseq, const uint32_t *const *r_starts,
                                 const uint32_t *const *r_ends, const size_t *r_n, const double *const *r_scores,
                                 const uint8_t *const *r_valid, const uint64_t *left_offsets, const uint32_t *ls,
                                 const uint32_t *le, const int *ops, int k, double *out, uint8_t *out_valid, char *err,
                                 size_t err_len);

/* ---- windows ----------------------------------------------------------- */

/* GRangesEmpty::from_windows -- src/granges.rs:634-666, generated on the device.  step 0 means
 * None (step = width).  The ranges stay on the device (pass grb_ranges_dev_* straight back as
 * ls/le of a query) until copied out. */
int grb_windows(grb_ctx *ctx, const uint32_t *seqlens, int nseq, uint32_t width, uint32_t step, int chop,
                grb_ranges **out);
size_t grb_ranges_len(const grb_ranges *r);
/* Per-seqname offsets into the window arrays (nseq+1 entries, host buffer). */
int grb_ranges_seq_offsets(const grb_ranges *r, uint64_t *offsets);
int grb_ranges_copy(const grb_ranges *r, uint32_t *seq_id, uint32_t *starts, uint32_t *ends);
const uint32_t *grb_ranges_dev_starts(const grb_ranges *r);
const uint32_t *grb_ranges_dev_ends(const grb_ranges *r);
void grb_ranges_destroy(grb_ranges *r);

/* ---- merging (`granges merge`, feature-density's per-feature merge) ---------- */

/* MergingEmptyIterator / MergingEmptyResultIterator / MergingResultIterator -- src/merging_iterators.rs:39-241,
 * as driven by Merge::run (src/commands.rs:616-673) and FeatureDensity (src/commands.rs:818-826).
 * The records are given in file order as `nruns` runs (run r = records [run_offsets[r], run_offsets[r+1]),
 * a maximal stretch of consecutive records of one seqname -- the iterator never merges across a seqname
 * change).  A record joins the open range iff distance_or_overlap(open, record) <= distance
 * (src/traits.rs:89-108: negative = minus the overlap width, 0 = book-ended, positive = gap) and extends its
 * end to max(end, record.end).  Start-sorted runs with distance >= 0 take a prefix-max scan; any other input
 * (negative distance, unsorted starts) is walked by the reference's own state machine, one warp per run.
 * starts / ends may be host or device pointers; run_offsets is a host array of nruns + 1 entries. */
int grb_merge_ranges(grb_ctx *ctx, const uint32_t *starts, const uint32_t *ends, const uint64_t *run_offsets, int nruns,
                     int64_t distance, grb_merged **out);
/* Number of merged ranges. */
size_t grb_merged_len(const grb_merged *m);
/* Copy out (any pointer may be NULL): the merged ranges, first[g] = input position of group g's first
 * record (len + 1 entries, first[len] = number of records; group g = records [first[g], first[g+1])), and
 * the merged-range offsets of the runs (nruns + 1 entries, host). */
int grb_merged_copy(const grb_merged *m, uint32_t *starts, uint32_t *ends, uint32_t *first, uint64_t *run_offsets_merged);
const uint32_t *grb_merged_dev_starts(const grb_merged *m);
const uint32_t *grb_merged_dev_ends(const grb_merged *m);
/* The closure of Merge::run for BED5 input (src/commands.rs:650-657): FloatOperation::run over the non-missing
 * scores of each merged group.  scores / valid are per input record (valid NULL = all Some); out / out_valid
 * are len x k, row-major.  Sums are the correctly rounded EXACT sums: a deliberate deviation from the
 * reference's left-to-right f64 fold, whose own rounding error (<= n * 2^-53 * sum|x|) is the only difference -- e.g.
 * 1e16 + 1 - 1e16 is 1 here and 0 there.  Indexes whose scores cannot be held in the fixed point (subnormals, more
 * than 126 bits of dynamic range: grb_index_exact_sums == 0) sum the members in index order instead, with that same
 * error bound.  min / max / median are bit-exact. */
int grb_merged_map_f64(grb_ctx *ctx, const grb_merged *m, const double *scores, const uint8_t *valid, const int *ops, int k,
                       double *out, uint8_t *out_valid);
void grb_merged_destroy(grb_merged *m);

/* ---- multi-GPU sharding (host-side, no device work) --------------------- */

/* One unit of independent work: lefts [left_begin, left_end) of seqname `seq`; every right
 * overlapping any of them has r.start < right_hi and r.end > right_lo.  Units never exchange
 * data (src/granges.rs:958-962 looks the right container up by the left's seqname). */
typedef struct grb_shard {
    int32_t seq;
    int32_t rank;
    uint64_t left_begin, left_end;
    uint32_t right_lo, right_hi;
} grb_shard;

/* Split nseq seqnames' lefts (sorted by start within a seqname; left_offsets as in the batch
 * calls, ls/le HOST arrays) into at most max_shards units of near-equal cost
 * (cost = lefts + weight_right * rights of the seqname, pro rata) and assign them to nranks
 * ranks in contiguous runs.  Returns the number of shards written. */
int grb_shard_plan(const uint64_t *left_offsets, const uint64_t *right_counts, int nseq, const uint32_t *ls,
                   const uint32_t *le, int nranks, grb_shard *shards, int max_shards);

#ifdef __cplusplus
}
#endif
#endif /* GRANGES_B200_H */
