"""O(NL*NR) numpy brute force of the overlap predicate and reductions.

TEST INFRASTRUCTURE ONLY.  Independent of the tree restatement in
granges_oracle.c; used to pin it on small random inputs.  Predicate:
`first <= end-1 && last >= start` with last = r.end-1 (ranges/coitrees.rs:48-57,
172-198) <=> r.start < l.end && r.end > l.start (book-ends do not overlap).
"""
from __future__ import annotations

import numpy as np


def overlap_matrix(ls, le, rs, re):
    ls = np.asarray(ls, np.int64)[:, None]
    le = np.asarray(le, np.int64)[:, None]
    rs = np.asarray(rs, np.int64)[None, :]
    re = np.asarray(re, np.int64)[None, :]
    return (rs < le) & (re > ls)


def counts(ls, le, rs, re):
    return overlap_matrix(ls, le, rs, re).sum(axis=1)


def groups(ls, le, rs, re, idx=None):
    """Per left: right data indices sorted by (start, end, index) (join.rs:121-128)."""
    m = overlap_matrix(ls, le, rs, re)
    rs = np.asarray(rs, np.int64)
    re = np.asarray(re, np.int64)
    idx = np.arange(len(rs), dtype=np.int64) if idx is None else np.asarray(idx, np.int64)
    order = np.lexsort((idx, re, rs))
    out = []
    for i in range(m.shape[0]):
        sel = order[m[i, order]]
        out.append(idx[sel])
    return out


def reduce_op(op, vals):
    """FloatOperation::run (data/operations.rs:53-109) -> (value, is_some)."""
    vals = np.asarray(vals, np.float64)
    n = len(vals)
    if op == "sum":
        s = 0.0
        for v in vals:
            s += v
        return s, True
    if op == "sum-not-empty":
        if n == 0:
            return 0.0, False
        s = 0.0
        for v in vals:
            s += v
        return s, True
    if op in ("min", "max"):
        f = vals[np.isfinite(vals)]
        if len(f) == 0:
            return 0.0, False
        return (f.min() if op == "min" else f.max()), True
    if op == "mean":
        if n == 0:
            return 0.0, False
        s = 0.0
        for v in vals:
            s += v
        return s / n, True
    if op == "median":
        if n == 0:
            return 0.0, False
        sv = np.sort(vals)
        mid = n // 2
        if n % 2 == 0:
            return (sv[mid - 1] + sv[mid]) / 2.0, True
        return sv[mid], True
    raise ValueError(op)


def map_joins(ls, le, rs, re, scores, valid, ops, idx=None):
    gs = groups(ls, le, rs, re, idx)
    ls = np.asarray(ls, np.int64)
    le = np.asarray(le, np.int64)
    idx_arr = np.arange(len(rs), dtype=np.int64) if idx is None else np.asarray(idx, np.int64)
    pos_of = {int(d): j for j, d in enumerate(idx_arr)}
    out = np.zeros((len(gs), len(ops)))
    ov = np.zeros((len(gs), len(ops)), np.uint8)
    for i, g in enumerate(gs):
        vals = [scores[d] for d in g if valid is None or valid[d]] if scores is not None else []
        for o, op in enumerate(ops):
            if op == "count":
                out[i, o], ov[i, o] = len(g), 1
            elif op in ("weighted-mean", "weighted-sum"):
                num = den = 0.0
                for d in g:
                    if valid is None or valid[d]:
                        j = pos_of[int(d)]
                        w = float(max(0, min(le[i], re[j]) - max(ls[i], rs[j])))
                        num += scores[d] * w
                        den += w
                out[i, o], ov[i, o] = (num if op == "weighted-sum" else (num / den if den > 0 else 0.0)), 1
            elif op == "overlap-bp":
                bp = 0
                for d in g:
                    j = pos_of[int(d)]
                    bp += max(0, min(le[i], re[j]) - max(ls[i], rs[j]))
                out[i, o], ov[i, o] = bp, 1
            else:
                v, ok = reduce_op(op, vals)
                out[i, o], ov[i, o] = (v if ok else 0.0), int(ok)
    return out, ov


def windows(seqlens, width, step=None, chop=False):
    """from_windows (granges.rs:634-666), literal loop."""
    out = []
    st = width if step is None else step
    for s, ln in enumerate(seqlens):
        start = 0
        while start < ln:
            end = start + width
            if end >= ln:
                if chop:
                    break
                out.append((s, start, min(end, ln)))
            else:
                out.append((s, start, end))
            start += st
    return out
