/*
 * granges_oracle.c -- CPU restatement of the reference's overlap-join hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under granges_b200/ (the product) may
 * link, load or call this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker / CPU baseline.
 *
 * What it restates (reference = vsbuffalo/granges v0.4.0, Rust):
 *   - COITrees::query / count_overlaps      src/ranges/coitrees.rs:48-64
 *   - half-open -> closed conversion          src/ranges/coitrees.rs:172-198
 *   - From<VecRanges> for COITrees (build)    src/ranges/coitrees.rs:78-84
 *   - left_overlaps (4 impls, same loop)      src/granges.rs:839-1034
 *   - LeftGroupedJoin + sort_ranges           src/join.rs:93-174
 *   - JoinDataLeftEmpty::map (gather)         src/join.rs:343-361
 *   - None-score flattening in granges_map    src/commands.rs:524-542
 *   - FloatOperation::run, median             src/data/operations.rs:17-109
 *   - filter_overlaps / antifilter_overlaps   src/granges.rs:1416-1609
 *   - GRangesEmpty::from_windows              src/granges.rs:634-666
 *   - GenericRange::overlap_width             src/traits.rs:73-80
 *
 * The interval arithmetic itself lives in the crates.io dependency
 * `coitrees = 0.4.0 [nosimd]` (Cargo.toml:16), whose source is NOT under
 * /root/reference and which cannot be built here (no cargo/rustc).  Its
 * published algorithm (BasicCOITree) is restated below: nodes sorted by
 * `first`, an implicit balanced binary tree over the sorted order with a
 * `subtree_last` augmentation, small subtrees (<= ORC_SIMPLE_SUBTREE nodes)
 * stored contiguously and scanned linearly, closed-interval queries.  The
 * one deliberate deviation is the node layout: coitrees stores the tree in
 * van-Emde-Boas order; this file stores it in pre-order (root, left subtree,
 * right subtree).  Layout changes cache behaviour only, never results; the
 * within-group visit order is unspecified in the reference (join.rs:97
 * "unsorted"), so groups are compared after the canonical (start,end,index)
 * sort (join.rs:121-128).
 *
 * Parity pinning: checked against every golden the reference's own tests
 * hold for this path (tests/golden/reference_goldens.json, see
 * tests/test_oracle_goldens.py) and against an O(NL*NR) brute force
 * (oracle/brute.py).  Large-scale parity with the real Rust binary is
 * unpinned in this container (no toolchain, no bedtools).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_SIMPLE_SUBTREE 32
#define ORC_NONE UINT32_MAX

enum {
    ORC_OP_SUM = 0,
    ORC_OP_SUM_NOT_EMPTY = 1,
    ORC_OP_MIN = 2,
    ORC_OP_MAX = 3,
    ORC_OP_MEAN = 4,
    ORC_OP_MEDIAN = 5,
    ORC_OP_COUNT = 6,      /* LeftGroupedJoin::num_overlaps, join.rs:153 */
    ORC_OP_OVERLAP_BP = 7, /* sum of overlap_widths, join.rs:158-163, commands.rs:839-843 */
    ORC_OP_WEIGHTED_MEAN = 8, /* examples/weighted_mean.rs:4-31: sum(score*width)/sum(width) over Some scores, else 0.0 */
    ORC_OP_WEIGHTED_SUM = 9,  /* the numerator of the same closure: score_sum (examples/weighted_mean.rs:17-21) */
};

/* One node = coitrees::IntervalNode<usize, usize> restated: closed interval
 * [first, last] (i32), metadata = data index. */
typedef struct {
    int32_t subtree_last;
    int32_t first;
    int32_t last;
    uint32_t left;  /* node index or ORC_NONE; for a simple subtree left == right == size */
    uint32_t right;
    uint64_t metadata;
} orc_node;

typedef struct orc_tree {
    orc_node *nodes;
    size_t n;
    uint32_t root;
} orc_tree;

typedef struct {
    int32_t first, last;
    uint64_t metadata;
} orc_iv;

static int cmp_iv(const void *a, const void *b) {
    const orc_iv *x = (const orc_iv *)a, *y = (const orc_iv *)b;
    if (x->first != y->first) return x->first < y->first ? -1 : 1;
    if (x->last != y->last) return x->last < y->last ? -1 : 1;
    if (x->metadata != y->metadata) return x->metadata < y->metadata ? -1 : 1;
    return 0;
}

/* Recursive layout of sorted[lo,hi) into nodes[*next...], returns node index. */
static uint32_t build_rec(const orc_iv *sorted, size_t lo, size_t hi, orc_node *nodes, uint32_t *next) {
    size_t size = hi - lo;
    if (size == 0) return ORC_NONE;
    uint32_t me = *next;
    if (size <= ORC_SIMPLE_SUBTREE) {
        /* simple subtree: contiguous, sorted, linearly scanned */
        int32_t sl = INT32_MIN;
        for (size_t i = lo; i < hi; i++)
            if (sorted[i].last > sl) sl = sorted[i].last;
        for (size_t i = lo; i < hi; i++) {
            orc_node *nd = &nodes[*next];
            nd->first = sorted[i].first;
            nd->last = sorted[i].last;
            nd->metadata = sorted[i].metadata;
            nd->subtree_last = sorted[i].last;
            nd->left = nd->right = 1;
            (*next)++;
        }
        nodes[me].subtree_last = sl;
        nodes[me].left = nodes[me].right = (uint32_t)size;
        return me;
    }
    size_t mid = lo + size / 2;
    (*next)++;
    orc_node *nd = &nodes[me];
    nd->first = sorted[mid].first;
    nd->last = sorted[mid].last;
    nd->metadata = sorted[mid].metadata;
    uint32_t l = build_rec(sorted, lo, mid, nodes, next);
    uint32_t r = build_rec(sorted, mid + 1, hi, nodes, next);
    nd = &nodes[me];
    nd->left = l;
    nd->right = r;
    int32_t sl = nd->last;
    if (l != ORC_NONE && nodes[l].subtree_last > sl) sl = nodes[l].subtree_last;
    if (r != ORC_NONE && nodes[r].subtree_last > sl) sl = nodes[r].subtree_last;
    nd->subtree_last = sl;
    /* a non-simple node must not look like a simple one (left == right) */
    return me;
}

/* COITrees::from(VecRanges) -- ranges/coitrees.rs:78-84, GenericInterval impls :172-198.
 * Returns NULL on invalid input: end <= start (RangeIndexed::new asserts, ranges/mod.rs:109)
 * or a coordinate that does not fit i32 (ranges/coitrees.rs:173-178 `try_into().unwrap()`). */
orc_tree *orc_tree_build(const uint32_t *starts, const uint32_t *ends, const uint64_t *idx, size_t n) {
    orc_tree *t = (orc_tree *)calloc(1, sizeof(orc_tree));
    if (!t) return NULL;
    t->n = n;
    t->root = ORC_NONE;
    if (n == 0) return t;
    orc_iv *sorted = (orc_iv *)malloc(n * sizeof(orc_iv));
    t->nodes = (orc_node *)malloc(n * sizeof(orc_node));
    if (!sorted || !t->nodes) goto fail;
    for (size_t i = 0; i < n; i++) {
        if (ends[i] <= starts[i] || ends[i] > (uint32_t)INT32_MAX) goto fail;
        sorted[i].first = (int32_t)starts[i];
        sorted[i].last = (int32_t)ends[i] - 1; /* right-inclusive "last" */
        sorted[i].metadata = idx ? idx[i] : (uint64_t)i;
    }
    qsort(sorted, n, sizeof(orc_iv), cmp_iv);
    uint32_t next = 0;
    t->root = build_rec(sorted, 0, n, t->nodes, &next);
    free(sorted);
    return t;
fail:
    free(sorted);
    free(t->nodes);
    free(t);
    return NULL;
}

void orc_tree_free(orc_tree *t) {
    if (!t) return;
    free(t->nodes);
    free(t);
}

size_t orc_tree_len(const orc_tree *t) { return t ? t->n : 0; }

typedef void (*orc_visit)(const orc_node *, void *);

static inline int iv_overlaps(int32_t fa, int32_t la, int32_t fb, int32_t lb) { return fa <= lb && la >= fb; }

/* BasicCOITree::query recursion restated (closed [first,last]). */
static void query_rec(const orc_node *nodes, uint32_t root, int32_t first, int32_t last, orc_visit visit, void *ud) {
    const orc_node *node = &nodes[root];
    if (node->left == node->right) {
        uint32_t sz = node->right;
        for (uint32_t k = root; k < root + sz; k++) {
            const orc_node *nd = &nodes[k];
            if (last < nd->first) break;
            if (first <= nd->last) visit(nd, ud);
        }
        return;
    }
    if (iv_overlaps(node->first, node->last, first, last)) visit(node, ud);
    if (node->left != ORC_NONE && nodes[node->left].subtree_last >= first)
        query_rec(nodes, node->left, first, last, visit, ud);
    if (node->right != ORC_NONE && iv_overlaps(node->first, nodes[node->right].subtree_last, first, last))
        query_rec(nodes, node->right, first, last, visit, ud);
}

/* COITrees::query -- ranges/coitrees.rs:48-57: query(first=start, last=end-1). */
static void tree_query(const orc_tree *t, uint32_t start, uint32_t end, orc_visit visit, void *ud) {
    if (!t || t->n == 0) return;
    query_rec(t->nodes, t->root, (int32_t)start, (int32_t)end - 1, visit, ud);
}

static void visit_count(const orc_node *nd, void *ud) {
    (void)nd;
    (*(size_t *)ud)++;
}

/* COITrees::count_overlaps -- ranges/coitrees.rs:60-64. */
size_t orc_count_overlaps(const orc_tree *t, uint32_t start, uint32_t end) {
    size_t c = 0;
    tree_query(t, start, end, visit_count, &c);
    return c;
}

/* RangeTuple (join.rs:17): (start, end, Some(index)). */
typedef struct {
    uint32_t start, end;
    uint64_t index;
} orc_rt;

typedef struct {
    orc_rt *v;
    size_t len, cap;
} orc_group;

static void group_push(orc_group *g, orc_rt r) {
    if (g->len == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 4;
        g->v = (orc_rt *)realloc(g->v, g->cap * sizeof(orc_rt));
    }
    g->v[g->len++] = r;
}

/* LeftGroupedJoin::add_right (join.rs:116-118) via IntervalNode -> GenericRange
 * (ranges/coitrees.rs:216-226: start = first, end = last + 1, index = metadata). */
static void visit_add_right(const orc_node *nd, void *ud) {
    orc_rt r = {(uint32_t)nd->first, (uint32_t)nd->last + 1u, nd->metadata};
    group_push((orc_group *)ud, r);
}

static int cmp_rt(const void *a, const void *b) {
    const orc_rt *x = (const orc_rt *)a, *y = (const orc_rt *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    if (x->index != y->index) return x->index < y->index ? -1 : 1;
    return 0;
}

/* One-range query, hits returned in canonical order.  Returns the number of hits;
 * writes at most cap of them. */
size_t orc_query(const orc_tree *t, uint32_t start, uint32_t end, uint64_t *out_idx, uint32_t *out_start,
                 uint32_t *out_end, size_t cap) {
    orc_group g = {0};
    tree_query(t, start, end, visit_add_right, &g);
    qsort(g.v, g.len, sizeof(orc_rt), cmp_rt);
    for (size_t i = 0; i < g.len && i < cap; i++) {
        if (out_idx) out_idx[i] = g.v[i].index;
        if (out_start) out_start[i] = g.v[i].start;
        if (out_end) out_end[i] = g.v[i].end;
    }
    size_t n = g.len;
    free(g.v);
    return n;
}

/* left_overlaps -- granges.rs:935-984 loop, CSR form of Vec<LeftGroupedJoin>.
 * Every left gets a group (left join), a missing right seqname (t == NULL)
 * gives empty groups.  Groups are sort_ranges()-ordered (join.rs:121-128).
 * Buffers (*idx, *rs, *re) are malloc'd; free with orc_free. */
int orc_left_overlaps(const orc_tree *t, const uint32_t *ls, const uint32_t *le, size_t nl, uint64_t *offsets,
                      uint64_t **idx, uint32_t **rs, uint32_t **re) {
    size_t cap = 1024, len = 0;
    uint64_t *oi = (uint64_t *)malloc(cap * sizeof(uint64_t));
    uint32_t *os = (uint32_t *)malloc(cap * sizeof(uint32_t));
    uint32_t *oe = (uint32_t *)malloc(cap * sizeof(uint32_t));
    orc_group g = {0};
    offsets[0] = 0;
    for (size_t i = 0; i < nl; i++) {
        g.len = 0;
        tree_query(t, ls[i], le[i], visit_add_right, &g);
        qsort(g.v, g.len, sizeof(orc_rt), cmp_rt);
        if (len + g.len > cap) {
            while (len + g.len > cap) cap *= 2;
            oi = (uint64_t *)realloc(oi, cap * sizeof(uint64_t));
            os = (uint32_t *)realloc(os, cap * sizeof(uint32_t));
            oe = (uint32_t *)realloc(oe, cap * sizeof(uint32_t));
        }
        for (size_t j = 0; j < g.len; j++) {
            oi[len] = g.v[j].index;
            os[len] = g.v[j].start;
            oe[len] = g.v[j].end;
            len++;
        }
        offsets[i + 1] = len;
    }
    free(g.v);
    *idx = oi;
    *rs = os;
    *re = oe;
    return 0;
}

void orc_free(void *p) { free(p); }

/* GenericRange::overlap_width -- traits.rs:73-80. */
uint32_t orc_overlap_width(uint32_t a_start, uint32_t a_end, uint32_t b_start, uint32_t b_end) {
    uint32_t s = a_start > b_start ? a_start : b_start;
    uint32_t e = a_end < b_end ? a_end : b_end;
    if (s >= e) return 0;
    return e - s;
}

/* --- data/operations.rs ------------------------------------------------- */

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* k-th order statistic.  The reference uses select_nth_unstable_by with
 * partial_cmp (NaN panics, operations.rs:22-29 -- callers must not pass NaN);
 * the k-th order statistic is unique as a value, so a full sort gives the
 * identical result. */
static double select_nth(double *v, size_t n, size_t k) {
    if (n <= 24) {
        for (size_t i = 1; i < n; i++) {
            double x = v[i];
            size_t j = i;
            while (j > 0 && v[j - 1] > x) {
                v[j] = v[j - 1];
                j--;
            }
            v[j] = x;
        }
    } else {
        qsort(v, n, sizeof(double), cmp_double);
    }
    return v[k];
}

/* median() -- data/operations.rs:17-32. Returns 0 if empty (None). */
int orc_median(double *v, size_t n, double *out) {
    if (n == 0) return 0;
    size_t mid = n / 2;
    if (n % 2 == 0) {
        double lower = select_nth(v, n, mid - 1);
        double upper = select_nth(v, n, mid);
        *out = (lower + upper) / 2.0;
    } else {
        *out = select_nth(v, n, mid);
    }
    return 1;
}

/* FloatOperation::run -- data/operations.rs:53-109.  Returns 1 for
 * DatumType::Float64(*out), 0 for DatumType::NoValue. */
int orc_float_op(int op, double *data, size_t n, double *out) {
    switch (op) {
    case ORC_OP_SUM: {
        double s = 0.0;
        for (size_t i = 0; i < n; i++) s += data[i];
        *out = s;
        return 1;
    }
    case ORC_OP_SUM_NOT_EMPTY: {
        if (n == 0) return 0;
        double s = 0.0;
        for (size_t i = 0; i < n; i++) s += data[i];
        *out = s;
        return 1;
    }
    case ORC_OP_MIN: {
        int have = 0;
        double m = 0;
        for (size_t i = 0; i < n; i++) {
            if (!isfinite(data[i])) continue;
            if (!have || data[i] < m) { /* min_by keeps the first of equal minima */
                m = data[i];
                have = 1;
            }
        }
        *out = m;
        return have;
    }
    case ORC_OP_MAX: {
        int have = 0;
        double m = 0;
        for (size_t i = 0; i < n; i++) {
            if (!isfinite(data[i])) continue;
            if (!have || data[i] >= m) { /* max_by keeps the last of equal maxima */
                m = data[i];
                have = 1;
            }
        }
        *out = m;
        return have;
    }
    case ORC_OP_MEAN: {
        if (n == 0) return 0;
        double s = 0.0;
        for (size_t i = 0; i < n; i++) s += data[i];
        *out = s / (double)n;
        return 1;
    }
    case ORC_OP_MEDIAN:
        return orc_median(data, n, out);
    default:
        return -1;
    }
}

/* left_overlaps + map_joins for ONE seqname, following the reference's
 * structure: per left a fresh group is built by the tree query's callback,
 * the right data is gathered by index (join.rs:353-356), None scores are
 * dropped (commands.rs:526-531), then each op scans the gathered slice
 * (commands.rs:534-541).  `valid` NULL => every score is Some.
 * t == NULL => the right has no such seqname: every group is empty.
 * COUNT / OVERLAP_BP are join-metadata reductions (join.rs:153-163) and see
 * every overlapping right, including None-score ones.
 * nthreads > 1 splits the lefts across POSIX threads (the reference itself is
 * single-threaded; this is the "generous" CPU bound). */
typedef struct {
    const orc_tree *t;
    const double *scores;
    const uint8_t *valid;
    const uint32_t *ls, *le;
    size_t nl;
    const int *ops;
    int k;
    double *out;
    uint8_t *out_valid;
    size_t *next; /* shared chunk cursor */
    int err;
} orc_map_job;

#define ORC_CHUNK 4096

static void *map_worker(void *arg) {
    orc_map_job *jb = (orc_map_job *)arg;
    orc_group g = {0};
    double *vals = NULL;
    size_t vcap = 0;
    for (;;) {
        size_t lo = __atomic_fetch_add(jb->next, (size_t)ORC_CHUNK, __ATOMIC_RELAXED);
        if (lo >= jb->nl) break;
        size_t hi = lo + ORC_CHUNK < jb->nl ? lo + ORC_CHUNK : jb->nl;
        for (size_t i = lo; i < hi; i++) {
            g.len = 0;
            tree_query(jb->t, jb->ls[i], jb->le[i], visit_add_right, &g);
            if (g.len > vcap) {
                vcap = g.len * 2;
                vals = (double *)realloc(vals, vcap * sizeof(double));
            }
            size_t nv = 0;
            uint64_t bp = 0;
            double score_sum = 0.0, total_width = 0.0;
            for (size_t j = 0; j < g.len; j++) {
                uint64_t di = g.v[j].index;
                uint32_t w = orc_overlap_width(g.v[j].start, g.v[j].end, jb->ls[i], jb->le[i]);
                bp += w;
                if (jb->scores && (!jb->valid || jb->valid[di])) {
                    vals[nv++] = jb->scores[di];
                    score_sum += jb->scores[di] * (double)w; /* examples/weighted_mean.rs:17-21 */
                    total_width += (double)w;
                }
            }
            for (int o = 0; o < jb->k; o++) {
                double r = 0.0;
                int ok;
                if (jb->ops[o] == ORC_OP_COUNT) {
                    r = (double)g.len;
                    ok = 1;
                } else if (jb->ops[o] == ORC_OP_OVERLAP_BP) {
                    r = (double)bp;
                    ok = 1;
                } else if (jb->ops[o] == ORC_OP_WEIGHTED_MEAN) {
                    r = total_width > 0.0 ? score_sum / total_width : 0.0; /* examples/weighted_mean.rs:26-30 */
                    ok = 1;
                } else if (jb->ops[o] == ORC_OP_WEIGHTED_SUM) {
                    r = score_sum;
                    ok = 1;
                } else {
                    ok = orc_float_op(jb->ops[o], vals, nv, &r);
                    if (ok < 0) {
                        jb->err = 1;
                        ok = 0;
                    }
                }
                jb->out[i * (size_t)jb->k + o] = ok ? r : 0.0;
                jb->out_valid[i * (size_t)jb->k + o] = (uint8_t)ok;
            }
        }
    }
    free(g.v);
    free(vals);
    return NULL;
}

int orc_map_joins_f64(const orc_tree *t, const double *scores, const uint8_t *valid, const uint32_t *ls,
                      const uint32_t *le, size_t nl, const int *ops, int k, double *out, uint8_t *out_valid,
                      int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    size_t next = 0;
    orc_map_job jobs[256];
    pthread_t th[256];
    for (int w = 0; w < nthreads; w++) {
        orc_map_job jb = {t, scores, valid, ls, le, nl, ops, k, out, out_valid, &next, 0};
        jobs[w] = jb;
    }
    if (nthreads == 1) {
        map_worker(&jobs[0]);
        return jobs[0].err ? -1 : 0;
    }
    for (int w = 0; w < nthreads; w++) pthread_create(&th[w], NULL, map_worker, &jobs[w]);
    int err = 0;
    for (int w = 0; w < nthreads; w++) {
        pthread_join(th[w], NULL);
        err |= jobs[w].err;
    }
    return err ? -1 : 0;
}

/* filter_overlaps / antifilter_overlaps for ONE seqname -- granges.rs:1416-1441,
 * 1485-1508, 1547-1609: keep = (count_overlaps > 0) XOR anti; if the right has
 * no such seqname (t == NULL) the semi-join drops and the anti-join keeps. */
int orc_filter_overlaps(const orc_tree *t, const uint32_t *ls, const uint32_t *le, size_t nl, int anti,
                        uint8_t *keep) {
    for (size_t i = 0; i < nl; i++) {
        if (!t) {
            keep[i] = anti ? 1 : 0;
            continue;
        }
        int has = orc_count_overlaps(t, ls[i], le[i]) > 0;
        keep[i] = (uint8_t)(has != (anti != 0));
    }
    return 0;
}

/* GRangesEmpty::from_windows -- granges.rs:634-666, loop restated literally
 * (including the quirk that without chop the loop keeps stepping and emits a
 * truncated window for every remaining start; pinned by granges.rs:1876-1907).
 * step == 0 means None (step = width).  Returns the number of windows; writes
 * at most cap. */
size_t orc_windows(const uint32_t *seqlens, int nseq, uint32_t width, uint32_t step, int chop, uint32_t *out_seq,
                   uint32_t *out_start, uint32_t *out_end, size_t cap) {
    size_t n = 0;
    uint64_t st = step ? step : width;
    for (int s = 0; s < nseq; s++) {
        uint64_t len = seqlens[s];
        uint64_t start = 0;
        while (start < len) {
            uint64_t end = start + width;
            if (end >= len) {
                if (chop) break;
                end = end < len ? end : len;
                if (n < cap) {
                    out_seq[n] = (uint32_t)s;
                    out_start[n] = (uint32_t)start;
                    out_end[n] = (uint32_t)end;
                }
                n++;
            } else {
                if (n < cap) {
                    out_seq[n] = (uint32_t)s;
                    out_start[n] = (uint32_t)start;
                    out_end[n] = (uint32_t)end;
                }
                n++;
            }
            start += st;
        }
    }
    return n;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}


/* --- merging_iterators.rs ------------------------------------------------ */

/* GenericRange::distance_or_overlap -- traits.rs:89-108 (self = a, other = b). */
static int64_t orc_distance_or_overlap(uint32_t as, uint32_t ae, uint32_t bs, uint32_t be) {
    if (ae > bs && as < be) return -(int64_t)orc_overlap_width(as, ae, bs, be);
    if (ae == bs || as == be) return 0;
    uint32_t d1 = bs > ae ? bs - ae : 0; /* saturating_sub */
    uint32_t d2 = as > be ? as - be : 0;
    return (int64_t)(d1 > d2 ? d1 : d2);
}

/* MergingEmptyResultIterator (merging_iterators.rs:72-136) and MergingResultIterator
 * (:138-241) restated as one streaming pass over the records of a file, in file order.
 * seq[i] = seqname id of record i.  scores NULL => the "empty" (BED3) iterator; otherwise
 * the closure of Merge::run for BED5 (commands.rs:650-657): the non-missing scores of the
 * accumulated records go through FloatOperation::run(op).  Outputs (each may be NULL) have
 * room for n entries; returns the number of merged ranges.  out_first[g] = position of the
 * first record of merged range g. */
size_t orc_merge(const uint32_t *seq, const uint32_t *starts, const uint32_t *ends, size_t n, int64_t min_distance,
                 const double *scores, const uint8_t *valid, int op, uint32_t *out_seq, uint32_t *out_start,
                 uint32_t *out_end, uint64_t *out_first, double *out_val, uint8_t *out_ok) {
    size_t m = 0;
    int have_last = 0;
    uint32_t last_seq = 0, last_start = 0, last_end = 0;
    size_t first = 0;
    double *acc = scores ? (double *)malloc((n ? n : 1) * sizeof(double)) : NULL; /* accumulated_data (Some scores only) */
    size_t nacc = 0;
    for (size_t i = 0; i <= n; i++) {
        int flush = 0;
        if (i < n) {
            if (have_last) {
                int on_same_chrom = last_seq == seq[i];
                if (on_same_chrom && orc_distance_or_overlap(last_start, last_end, starts[i], ends[i]) <= min_distance) {
                    if (ends[i] > last_end) last_end = ends[i]; /* last_range.end = max(..) */
                    if (scores && (!valid || valid[i])) acc[nacc++] = scores[i];
                    continue;
                }
                flush = 1;
            }
        } else {
            flush = have_last;
        }
        if (flush) {
            if (out_seq) out_seq[m] = last_seq;
            if (out_start) out_start[m] = last_start;
            if (out_end) out_end[m] = last_end;
            if (out_first) out_first[m] = first;
            if (scores) {
                double v = 0.0;
                int ok = orc_float_op(op, acc, nacc, &v);
                if (out_val) out_val[m] = ok == 1 ? v : 0.0;
                if (out_ok) out_ok[m] = ok == 1;
            }
            m++;
            nacc = 0;
        }
        if (i < n) {
            have_last = 1;
            last_seq = seq[i];
            last_start = starts[i];
            last_end = ends[i];
            first = i;
            if (scores && (!valid || valid[i])) acc[nacc++] = scores[i];
        }
    }
    free(acc);
    return m;
}
