"""Two implementations of the same per-seqname interface, so every golden /
parity test body runs unchanged against the CPU oracle (-m "not gpu") and the
CUDA library through its C ABI (-m gpu)."""
from __future__ import annotations

import numpy as np


class OracleEngine:
    name = "oracle"

    def __init__(self):
        from oracle import oracle as orc

        self.orc = orc

    def build(self, starts, ends, idx=None, scores=None, valid=None):
        t = self.orc.Tree(starts, ends, idx)
        t.scores = None if scores is None else np.asarray(scores, np.float64)
        t.valid = None if valid is None else np.asarray(valid, np.uint8)
        return t

    def count(self, t, ls, le):
        return np.array([t.count_overlaps(int(a), int(b)) for a, b in zip(ls, le)], dtype=np.uint32)

    def filter(self, t, ls, le, anti=False):
        return self.orc.filter_overlaps(t, ls, le, anti)

    def left_overlaps(self, t, ls, le):
        return self.orc.left_overlaps(t, ls, le)

    def map_joins(self, t, ls, le, ops):
        if t is None:
            return self.orc.map_joins(None, None, None, ls, le, ops)
        return self.orc.map_joins(t, t.scores, t.valid, ls, le, ops)

    def windows(self, seqlens, width, step=None, chop=False):
        return self.orc.windows(seqlens, width, step, chop)


class CudaEngine:
    name = "cuda"

    def __init__(self):
        import __graft_entry__ as g

        g.build()  # no-op when the in-tree library is up to date
        import granges_b200 as gb

        self.gb = gb
        self.ctx = gb.Context(0)

    def build(self, starts, ends, idx=None, scores=None, valid=None):
        ix = self.ctx.index_build(starts, ends, idx)
        if scores is not None:
            ix.set_scores(scores, valid)
        return ix

    def count(self, t, ls, le):
        return self.ctx.count_overlaps(t, ls, le)

    def filter(self, t, ls, le, anti=False):
        return self.ctx.filter_overlaps(t, ls, le, anti)

    def left_overlaps(self, t, ls, le):
        return self.ctx.left_overlaps(t, ls, le)

    def map_joins(self, t, ls, le, ops):
        return self.ctx.map_joins(t, ls, le, ops)

    def windows(self, seqlens, width, step=None, chop=False):
        return self.ctx.windows(seqlens, width, step, chop)


def make_engine(name):
    return OracleEngine() if name == "oracle" else CudaEngine()
