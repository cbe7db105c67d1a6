"""The reference's container surface for the overlap-join path, over the C ABI.

Same names and argument meaning as vsbuffalo/granges v0.4.0 (paths relative to its `src/`), so the
parity tests in tests/test_reference_api.py read like the reference's own unit tests:

    GRangesEmpty::new_vec / GRanges::new_vec          granges.rs:286-295, 718-729
    push_range                                        granges.rs:731-755, 1211-1249
    GRangesEmpty::from_windows                        granges.rs:634-666
    into_coitrees                                     granges.rs:1365-1402
    map_data                                          granges.rs:757-775
    left_overlaps (4 impls) -> joins                  granges.rs:839-1034
    map_joins                                         granges.rs:1036-1191   (FloatOperation names, or a host closure)
    filter_overlaps / antifilter_overlaps             granges.rs:1404-1610
    LeftGroupedJoin::num_overlaps / overlap_widths    join.rs:153-163

Everything numeric runs on the GPU through `granges_b200.Context`; this module only keeps the per-seqname
vectors the reference keeps in its `GenomeMap` and re-assembles results in the reference's row order
(seqnames in genome order, then insertion order -- iterators.rs:43-76).  A closure passed to `map_joins`
runs on the host over the materialised pairs (`grb_left_overlaps`), exactly where the reference would run it.
"""
from __future__ import annotations

import numpy as np

from .api import Context, GRangesError

_ctx = None


def default_context():
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


class _Seq:
    __slots__ = ("starts", "ends", "index")

    def __init__(self):
        self.starts, self.ends, self.index = [], [], []


class GRangesEmpty:
    """`GRangesEmpty<VecRangesEmpty>` / `GRangesEmpty<COITreesEmpty>` (after into_coitrees)."""

    def __init__(self, seqlens, ctx=None):
        self.seqlens = dict(seqlens)
        self.ctx = ctx or default_context()
        self._seqs = {name: _Seq() for name in self.seqlens}
        self._n = 0
        self._trees = None  # per-seqname Index after into_coitrees()

    # -- constructors
    @classmethod
    def new_vec(cls, seqlens, ctx=None):
        return cls(seqlens, ctx)

    @classmethod
    def from_windows(cls, seqlens, width, step=None, chop=False, ctx=None):
        """granges.rs:634-666 -- generated on the device (grb_windows)."""
        gr = cls(seqlens, ctx)
        names = list(gr.seqlens)
        seq, s, e = gr.ctx.windows([gr.seqlens[n] for n in names], width, step, chop)
        for k, name in enumerate(names):
            m = seq == k
            sq = gr._seqs[name]
            sq.starts, sq.ends = [int(x) for x in s[m]], [int(x) for x in e[m]]
            sq.index = list(range(gr._n, gr._n + int(m.sum())))
            gr._n += int(m.sum())
        return gr

    # -- building
    def _check(self, seqname, start, end):
        if seqname not in self._seqs:
            raise GRangesError(-1, f"MissingSequence: The sequence name '{seqname}' is not found within the provided ranges container. "
                                   "Check the sequence names for typos or missing entries.")  # GRangesError::MissingSequence
        if not end > start:
            raise AssertionError("end > start")  # RangeEmpty::new / RangeIndexed::new assert (ranges/mod.rs:30,109)

    def push_range(self, seqname, start, end):
        self._check(seqname, start, end)
        sq = self._seqs[seqname]
        sq.starts.append(int(start))
        sq.ends.append(int(end))
        sq.index.append(self._n)
        self._n += 1
        self._trees = None

    def __len__(self):
        return self._n

    def is_empty(self):
        return self._n == 0

    def iter_ranges(self):
        """(seqname, start, end) in the reference's row order."""
        for name, sq in self._seqs.items():
            for s, e in zip(sq.starts, sq.ends):
                yield name, s, e

    # -- index
    def into_coitrees(self):
        self._trees = {}
        for name, sq in self._seqs.items():
            if sq.starts:  # a seqname without ranges has no container (granges.rs:1372-1401 iterates the populated ones)
                self._trees[name] = self._build(sq)
        return self

    def _build(self, sq):
        return self.ctx.index_build(np.array(sq.starts, np.uint32), np.array(sq.ends, np.uint32))

    def _tree(self, name):
        if self._trees is None:
            raise GRangesError(-1, "the right-hand ranges are not indexed: call into_coitrees() first")
        return self._trees.get(name)

    # -- joins
    def left_overlaps(self, right):
        return LeftJoin(self, right)

    def _filter(self, right, anti):
        out = type(self)(self.seqlens, self.ctx)
        for name, sq in self._seqs.items():
            if not sq.starts:
                continue
            tree = right._tree(name)
            # no container for this seqname on the right: the semi-join drops the lefts, the anti-join keeps them
            # (granges.rs:1433-1437, 1580-1593) -- which is what a NULL index answers
            keep = self.ctx.filter_overlaps(tree, np.array(sq.starts, np.uint32), np.array(sq.ends, np.uint32), anti)
            for k in np.flatnonzero(keep):
                out._push_from(self, name, int(k))
        return out

    def _push_from(self, src, name, k):
        sq = src._seqs[name]
        self.push_range(name, sq.starts[k], sq.ends[k])

    def filter_overlaps(self, right):
        return self._filter(right, False)

    def antifilter_overlaps(self, right):
        return self._filter(right, True)


class GRanges(GRangesEmpty):
    """`GRanges<VecRangesIndexed, Vec<U>>` with U = Option<f64> once `map_data` has extracted the scores."""

    def __init__(self, seqlens, ctx=None):
        super().__init__(seqlens, ctx)
        self.data = []

    def push_range(self, seqname, start, end, data=None):
        self._check(seqname, start, end)
        sq = self._seqs[seqname]
        sq.starts.append(int(start))
        sq.ends.append(int(end))
        sq.index.append(len(self.data))  # index = data.len() (granges.rs:733-754)
        self.data.append(data)
        self._n += 1
        self._trees = None

    def _push_from(self, src, name, k):
        sq = src._seqs[name]
        self.push_range(name, sq.starts[k], sq.ends[k], src.data[sq.index[k]])

    def map_data(self, func):
        self.data = [func(d) for d in self.data]
        self._trees = None
        return self

    def _build(self, sq):
        ix = self.ctx.index_build(np.array(sq.starts, np.uint32), np.array(sq.ends, np.uint32), np.array(sq.index, np.uint64))
        if all(d is None or isinstance(d, (int, float)) for d in self.data):
            scores = np.array([0.0 if d is None else float(d) for d in self.data], np.float64)
            valid = np.array([d is not None for d in self.data], np.uint8)
            ix.set_scores(scores, None if valid.all() else valid)
        return ix


class LeftJoin:
    """The result of `left_overlaps`: one LeftGroupedJoin per left range, in the reference's row order."""

    def __init__(self, left, right):
        self.left, self.right = left, right

    def __len__(self):
        return len(self.left)

    def _per_seq(self):
        for name, sq in self.left._seqs.items():
            if sq.starts:
                yield name, sq, self.right._tree(name), np.array(sq.starts, np.uint32), np.array(sq.ends, np.uint32)

    def num_overlaps(self):
        out = []
        for _, _, tree, ls, le in self._per_seq():
            out += [int(c) for c in self.left.ctx.count_overlaps(tree, ls, le)]
        return out

    def overlap_widths(self):
        out = []
        for _, _, tree, ls, le in self._per_seq():
            off, _, _, _, w = self.left.ctx.left_overlaps(tree, ls, le, widths=True)
            out += [[int(x) for x in w[int(off[i]):int(off[i + 1])]] for i in range(len(ls))]
        return out

    def map_joins(self, func):
        """`func`: a FloatOperation name ("sum", "mean", ...), a list of them (-> one row per left, None = NoValue),
        or a closure `f(left_range, right_data_list)` run on the host like the reference runs its closures."""
        ctx = self.left.ctx
        rows = []
        if callable(func):
            for name, sq, tree, ls, le in self._per_seq():
                off, ridx, _, _ = ctx.left_overlaps(tree, ls, le)
                data = getattr(self.right, "data", None)
                for i in range(len(ls)):
                    members = [int(j) for j in ridx[int(off[i]):int(off[i + 1])]]
                    rows.append(func((name, int(ls[i]), int(le[i])), [data[j] for j in members] if data is not None else members))
            return rows
        ops = [func] if isinstance(func, str) else list(func)
        for _, _, tree, ls, le in self._per_seq():
            out, ok = ctx.map_joins(tree, ls, le, ops)
            for i in range(len(ls)):
                vals = [float(out[i, c]) if ok[i, c] else None for c in range(len(ops))]
                rows.append(vals[0] if isinstance(func, str) else vals)
        return rows
