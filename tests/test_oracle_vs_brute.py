"""The tree restatement (oracle/granges_oracle.c) against the independent O(NL*NR)
numpy brute force (oracle/brute.py) on seeded random inputs, including the sizes
around the simple-subtree cutoff, ties, None scores and long intervals."""
import numpy as np
import pytest

from oracle import brute
from oracle import oracle as orc
from tests.randcases import ALL_OPS, random_case

SIZES = [(1, 1), (7, 0), (0, 5), (40, 31), (40, 32), (40, 33), (64, 65), (200, 100), (300, 1000), (150, 2500)]


@pytest.mark.parametrize("nl,nr", SIZES)
@pytest.mark.parametrize("kw", [dict(), dict(null_frac=0.3, score_kind="int"), dict(long_frac=0.05, score_kind="mixed"),
                                dict(maxlen=3, span=50)], ids=["plain", "nulls", "long", "dense-ties"])
def test_tree_vs_brute(nl, nr, kw):
    c = random_case(1000 + nl * 7 + nr, nl, nr, **kw)
    t = orc.Tree(c["rs"], c["re"])
    # counts
    want = brute.counts(c["ls"], c["le"], c["rs"], c["re"])
    got = np.array([t.count_overlaps(int(a), int(b)) for a, b in zip(c["ls"], c["le"])])
    assert (got == want).all()
    # groups (canonical order)
    off, idx, prs, pre = orc.left_overlaps(t, c["ls"], c["le"])
    gs = brute.groups(c["ls"], c["le"], c["rs"], c["re"])
    assert len(off) == nl + 1
    for i, g in enumerate(gs):
        seg = idx[int(off[i]):int(off[i + 1])]
        assert list(seg) == list(g)
        assert (prs[int(off[i]):int(off[i + 1])] == c["rs"][g]).all()
        assert (pre[int(off[i]):int(off[i + 1])] == c["re"][g]).all()
    # reductions
    out, ov = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ALL_OPS)
    bout, bov = brute.map_joins(c["ls"], c["le"], c["rs"], c["re"], c["scores"], c["valid"], ALL_OPS)
    assert (ov == bov).all()
    for o, op in enumerate(ALL_OPS):
        m = bov[:, o] == 1
        if op in ("sum", "sum-not-empty", "mean", "weighted-mean", "weighted-sum"):
            # visit order differs between tree and brute force: tolerance relative to sum |x|
            assert np.allclose(out[m, o], bout[m, o], rtol=1e-12, atol=1e-9 * (np.abs(c["scores"]).max() if nr else 1) * (1e3 if op == "weighted-sum" else 1))
        else:
            assert (out[m, o] == bout[m, o]).all(), op


def test_threads_agree():
    c = random_case(5, 20000, 3000, span=100000)
    t = orc.Tree(c["rs"], c["re"])
    a = orc.map_joins(t, c["scores"], None, c["ls"], c["le"], ["mean", "median", "count"], nthreads=1)
    b = orc.map_joins(t, c["scores"], None, c["ls"], c["le"], ["mean", "median", "count"], nthreads=4)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_windows_vs_brute():
    for lens, w, st, chop in [([35], 10, None, True), ([21, 13], 10, 2, False), ([1000, 999, 1], 100, 30, True),
                              ([5], 10, None, False), ([5], 10, None, True), ([100], 10, 10, True), ([100], 10, 10, False)]:
        seq, s, e = orc.windows(lens, w, st, chop)
        assert [(int(a), int(b), int(c)) for a, b, c in zip(seq, s, e)] == brute.windows(lens, w, st, chop)


def test_custom_data_index():
    c = random_case(9, 50, 80)
    idx = np.random.default_rng(1).permutation(80).astype(np.uint64)
    t = orc.Tree(c["rs"], c["re"], idx)
    off, got_idx, _, _ = orc.left_overlaps(t, c["ls"], c["le"])
    gs = brute.groups(c["ls"], c["le"], c["rs"], c["re"], idx)
    for i, g in enumerate(gs):
        assert list(got_idx[int(off[i]):int(off[i + 1])]) == list(g)
