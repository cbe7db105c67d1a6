"""Medians from the per-run wavelet trees of k_sweep3 (csrc/sweep3.cuh) against the oracle's member-by-member median
(src/data/operations.rs:17-32, 89-97) and against the member sweep they replace: bit-identical on every shape, including
the ones the default run-length rule would leave to the sweep (GRB_WAVE_RATIO forces the trees wherever the slice fits)."""
import numpy as np
import pytest

from tests.randcases import random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb():
    import __graft_entry__ as g

    g.build()
    import granges_b200 as gb

    return gb


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle as o

    o.build()
    return o


def contexts(gb, monkeypatch):
    monkeypatch.setenv("GRB_WAVE_RATIO", "1000000")
    always = gb.Context(0)
    monkeypatch.setenv("GRB_WAVE_RATIO", "0")
    never = gb.Context(0)
    monkeypatch.delenv("GRB_WAVE_RATIO")
    return always, never, gb.Context(0)


def check(ctxs, orc, case, ops):
    t = orc.Tree(case["rs"], case["re"])
    want, wv = orc.map_joins(t, case["scores"], case["valid"], case["ls"], case["le"], ops)
    for ctx in ctxs:
        ix = ctx.index_build(case["rs"], case["re"]).set_scores(case["scores"], case["valid"])
        out, ov = ctx.map_joins(ix, case["ls"], case["le"], ops)
        assert (ov == wv).all(), "validity differs"
        for c, op in enumerate(ops):
            m = wv[:, c] == 1
            if op in ("min", "max", "median", "count"):
                assert (out[m, c] == want[m, c]).all(), (op, int((out[m, c] != want[m, c]).sum()), int(m.sum()))
        ix.close()


SHAPES = [
    # nl, nr, span, maxlen, sorted_left
    (1, 1, 50, 10, True),
    (3000, 3000, 100_000, 900, True),      # ~ 27 members per left, one or two runs per tile
    (5000, 1500, 20_000, 400, True),       # dense lefts, groups of ~60
    (1500, 6000, 20_000, 300, True),       # slices beyond 2048 rights per 1024 lefts: the runs are halved
    (4000, 4000, 3000, 200, True),         # heavy coverage: groups of several hundred
    (2500, 2500, 60_000, 500, False),      # unsorted lefts: slices that rarely fit
    (2000, 300, 40_000, 50, True),         # many empty groups
    (1024, 2048, 4096, 64, True),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind,null_frac", [("uniform", 0.0), ("int", 0.0), ("uniform", 0.3), ("int", 0.9)])
def test_wavelet_median_vs_oracle(gb, orc, monkeypatch, shape, kind, null_frac):
    nl, nr, span, maxlen, sorted_left = shape
    ctxs = contexts(gb, monkeypatch)
    case = random_case((nl * 7 + nr + len(kind)) % 1000, nl, nr, span=span, maxlen=maxlen, null_frac=null_frac, score_kind=kind,
                       sorted_left=sorted_left)
    check(ctxs, orc, case, ["median"])
    check(ctxs, orc, case, ["min", "median", "max", "mean", "count"])


def test_wavelet_median_long_rights(gb, orc, monkeypatch):
    """A few very long rights among short ones: far reach-back, large C / R parts of the B sequence."""
    ctxs = contexts(gb, monkeypatch)
    case = random_case(11, 6000, 6000, span=150_000, maxlen=300, long_frac=0.002, sorted_left=True)
    check(ctxs, orc, case, ["median", "min", "max"])


def test_wavelet_all_equal_scores(gb, orc, monkeypatch):
    ctxs = contexts(gb, monkeypatch)
    case = random_case(12, 3000, 3000, span=50_000, maxlen=500, sorted_left=True)
    case["scores"] = np.full(3000, 2.5)
    check(ctxs, orc, case, ["median", "min", "max"])


def test_wavelet_batch_of_seqnames(gb, orc, monkeypatch):
    """Several seqnames in one call: tiles of different indexes, a seqname without rights, ragged tile edges."""
    always, never, default = contexts(gb, monkeypatch)
    rng = np.random.default_rng(3)
    cases = [random_case(20 + s, n, m, span=80_000, maxlen=700, sorted_left=True) for s, (n, m) in enumerate([(2500, 2000), (700, 0), (1800, 3000)])]
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    off = np.cumsum([0] + [len(c["ls"]) for c in cases]).astype(np.uint64)
    ops = ["median", "max"]
    want = []
    for c in cases:
        if len(c["rs"]):
            w, v = orc.map_joins(orc.Tree(c["rs"], c["re"]), c["scores"], None, c["ls"], c["le"], ops)
        else:
            w, v = orc.map_joins(None, None, None, c["ls"], c["le"], ops)
        want.append((w, v))
    ww = np.concatenate([w for w, _ in want])
    wv = np.concatenate([v for _, v in want])
    for ctx in (always, never, default):
        idxs = [ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"]) if len(c["rs"]) else None for c in cases]
        out, ov = ctx.map_joins_batch(idxs, off, ls, le, ops)
        assert (ov == wv).all()
        m = wv == 1
        assert (out[m] == ww[m]).all()
