"""The oracle's restatement of the merging iterators (src/merging_iterators.rs) against the reference's own
known answers and against an independent pure-Python replay of the streaming rule."""
import numpy as np
import pytest

from tests.mergecases import merge_case, py_merge


def _records(G):
    recs = G["merging_iterators"]["records"]
    names = sorted({r[0] for r in recs})
    seq = np.array([names.index(r[0]) for r in recs], np.uint32)
    return names, seq, np.array([r[1] for r in recs], np.uint32), np.array([r[2] for r in recs], np.uint32), np.array([r[3] for r in recs])


def test_merge_goldens(goldens):
    from oracle import oracle as orc

    names, seq, s, e, sc = _records(goldens)
    for case in goldens["merging_iterators"]["cases"]:
        if "op" in case:
            mseq, ms, me, first, val, ok = orc.merge(seq, s, e, case["distance"], sc, None, case["op"])
            got = [[names[a], int(b), int(c), float(v)] for a, b, c, v in zip(mseq, ms, me, val)]
            assert ok.all()
        else:
            mseq, ms, me, first = orc.merge(seq, s, e, case["distance"])
            got = [[names[a], int(b), int(c)] for a, b, c in zip(mseq, ms, me)]
        assert got == case["merged"], case["cite"]  # sums included: the oracle adds left to right like the reference


@pytest.mark.parametrize("distance", [0, 1, 25, -1, -7, -60])
@pytest.mark.parametrize("kind", ["sorted", "unsorted", "dense", "sparse"])
def test_merge_vs_python_replay(kind, distance):
    from oracle import oracle as orc

    c = merge_case(131 + distance, kind, 3000)
    for op in (None, "sum", "median", "min"):
        want = py_merge(c["seq"], c["starts"], c["ends"], distance, c["scores"] if op else None, c["valid"], op)
        got = orc.merge(c["seq"], c["starts"], c["ends"], distance, c["scores"] if op else None, c["valid"], op)
        assert len(got[0]) == len(want["start"])
        assert (got[0] == want["seq"]).all() and (got[1] == want["start"]).all() and (got[2] == want["end"]).all()
        assert (got[3] == want["first"]).all()
        if op:
            assert (got[5] == want["ok"]).all()
            m = want["ok"] == 1
            assert (got[4][m] == want["val"][m]).all()
