"""Seeded record streams for the merge tests and an independent pure-Python replay of the reference's
streaming rule (src/merging_iterators.rs:138-241, src/traits.rs:89-108)."""
import numpy as np


def merge_case(seed, kind, n, nseq=4):
    rng = np.random.default_rng(seed)
    span = {"dense": 4 * n, "sparse": 400 * n}.get(kind, 40 * n)
    seq = np.sort(rng.integers(0, nseq, n)).astype(np.uint32)
    if kind == "unsorted":  # seqnames come back, starts are in any order
        seq = rng.integers(0, nseq, n).astype(np.uint32)
        seq = np.repeat(seq[: max(1, n // 50)], 50)[:n]
    ln = rng.integers(1, 120, n)
    st = rng.integers(0, span, n)
    if kind != "unsorted":
        o = np.lexsort((st, seq))
        seq, st, ln = seq[o], st[o], ln[o]
    sc = np.round(rng.standard_normal(n) * 100) / 16
    valid = (rng.random(n) > 0.15).astype(np.uint8)
    return dict(seq=seq, starts=st.astype(np.uint32), ends=(st + ln).astype(np.uint32), scores=sc, valid=valid)


def run_offsets(seq):
    """Maximal stretches of consecutive records with the same seqname."""
    seq = np.asarray(seq)
    if len(seq) == 0:
        return np.array([0], np.uint64)
    cuts = np.flatnonzero(seq[1:] != seq[:-1]) + 1
    return np.concatenate([[0], cuts, [len(seq)]]).astype(np.uint64)


def _doo(a_s, a_e, b_s, b_e):
    if a_e > b_s and a_s < b_e:
        return -(min(a_e, b_e) - max(a_s, b_s))
    if a_e == b_s or a_s == b_e:
        return 0
    return max(max(b_s - a_e, 0), max(a_s - b_e, 0))


def _op(op, xs):
    import math

    if op == "sum":
        s = 0.0
        for x in xs:
            s += x
        return s, 1
    if not xs:
        return 0.0, 0
    if op == "min":
        f = [x for x in xs if math.isfinite(x)]
        return (min(f), 1) if f else (0.0, 0)
    if op == "median":
        v = sorted(xs)
        k = len(v) // 2
        return (v[k], 1) if len(v) % 2 else ((v[k - 1] + v[k]) / 2.0, 1)
    raise ValueError(op)


def py_merge(seq, starts, ends, distance, scores=None, valid=None, op=None):
    out = dict(seq=[], start=[], end=[], first=[], val=[], ok=[])
    last = None
    acc = []
    for i in range(len(starts)):
        q, s, e = int(seq[i]), int(starts[i]), int(ends[i])
        if last is not None and last[0] == q and _doo(last[1], last[2], s, e) <= distance:
            last[2] = max(last[2], e)
        else:
            if last is not None:
                _emit(out, last, acc, op)
            last = [q, s, e, i]
            acc = []
        if scores is not None and (valid is None or valid[i]):
            acc.append(float(scores[i]))
    if last is not None:
        _emit(out, last, acc, op)
    return {k: np.array(v) for k, v in out.items()}


def _emit(out, last, acc, op):
    out["seq"].append(last[0])
    out["start"].append(last[1])
    out["end"].append(last[2])
    out["first"].append(last[3])
    if op:
        v, ok = _op(op, acc)
        out["val"].append(v)
        out["ok"].append(ok)
