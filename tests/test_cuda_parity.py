"""Parity of the CUDA path (through the C ABI) against the CPU oracle: the reference's golden
vectors, seeded random cases at sizes the oracle finishes in seconds, edge cases, and
size-independent properties at larger sizes.  Bit-exact for pairs, counts, min, max, median;
sum / mean within 1e-12 of sum|x| over the group (f64)."""
import numpy as np
import pytest

from tests import golden_checks as gc
from tests.mapchecks import SUM_TOL, assert_map_equal, finite_abs
from tests.randcases import ALL_OPS, random_case

pytestmark = pytest.mark.gpu



@pytest.fixture(scope="module")
def eng():
    from tests.engines import CudaEngine

    return CudaEngine()


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


@pytest.mark.parametrize("check", gc.ALL_ENGINE_CHECKS, ids=lambda f: f.__name__)
def test_golden(eng, goldens, check):
    check(eng, goldens)


def run_case(eng, orc, c, ops=ALL_OPS, idx=None):
    t = orc.Tree(c["rs"], c["re"], idx)
    ix = eng.build(c["rs"], c["re"], idx, c["scores"], c["valid"])
    # counts + filter
    want_counts = np.array([t.count_overlaps(int(a), int(b)) for a, b in zip(c["ls"], c["le"])], dtype=np.uint32) \
        if len(c["ls"]) <= 5000 else orc.map_joins(t, None, None, c["ls"], c["le"], ["count"])[0][:, 0].astype(np.uint32)
    got_counts = eng.count(ix, c["ls"], c["le"])
    assert (got_counts == want_counts).all()
    assert (eng.filter(ix, c["ls"], c["le"]) == (want_counts > 0)).all()
    assert (eng.filter(ix, c["ls"], c["le"], anti=True) == (want_counts == 0)).all()
    # reductions
    want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    abs_sums = orc.map_joins(t, finite_abs(c["scores"]), c["valid"], c["ls"], c["le"], ["sum"])[0][:, 0]
    nvalid = orc.map_joins(t, np.ones(len(c["rs"])), c["valid"], c["ls"], c["le"], ["sum"])[0][:, 0]
    out, ov = eng.map_joins(ix, c["ls"], c["le"], ops)
    assert_map_equal(out, ov, want, wv, ops, abs_sums, nvalid)
    return ix, t


SIZES = [(1, 1), (7, 0), (0, 5), (40, 31), (40, 33), (64, 65), (300, 1000), (1500, 2500), (5000, 300), (1025, 4097)]
KINDS = {
    "plain": dict(),
    "nulls-int": dict(null_frac=0.3, score_kind="int"),
    "long-mixed": dict(long_frac=0.05, score_kind="mixed"),
    "dense-ties": dict(maxlen=3, span=50),
    "wide": dict(score_kind="wide"),
    "nonfinite": dict(score_kind="nonfinite"),
    "decimal": dict(score_kind="decimal"),
    "signed-decimal-nulls": dict(score_kind="signed-decimal", null_frac=0.2),
    "sorted-left": dict(sorted_left=True, span=20000),
    "nulls-sorted": dict(null_frac=0.5, sorted_left=True, span=5000),
}


@pytest.mark.parametrize("nl,nr", SIZES)
@pytest.mark.parametrize("kind", list(KINDS))
def test_random_vs_oracle(eng, orc, nl, nr, kind):
    c = random_case(2000 + nl * 7 + nr, nl, nr, **KINDS[kind])
    run_case(eng, orc, c)


def test_sum_paths_agree(eng, orc, monkeypatch):
    """SUM / MEAN served by compact prefixes, wide prefixes and the sweep must all meet the bound;
    compact and wide are both exactly-rounded sums, so they agree bit for bit."""
    c = random_case(77, 3000, 4000, span=30000, sorted_left=True)
    res = {}
    for mode in ("1", "2", "0"):
        monkeypatch.setenv("GRB_SUM_MODE", mode)
        ix = eng.build(c["rs"], c["re"], None, c["scores"], None)
        assert ix.exact_sums == (mode != "0")
        res[mode] = eng.map_joins(ix, c["ls"], c["le"], ["sum", "mean", "sum-not-empty"])
    assert (res["1"][0] == res["2"][0]).all() and (res["1"][1] == res["2"][1]).all()
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], None, c["ls"], c["le"], ["sum", "mean", "sum-not-empty"])
    for mode in res:
        assert (res[mode][1] == wv).all()
        assert np.allclose(res[mode][0], want, rtol=1e-12, atol=0)


def test_exact_sum_is_correctly_rounded(eng):
    """The prefix path returns the correctly rounded value of the exact sum (math.fsum)."""
    import math

    c = random_case(5, 400, 600, span=3000, score_kind="mixed")
    ix = eng.build(c["rs"], c["re"], None, c["scores"], None)
    assert ix.exact_sums
    out, ov = eng.map_joins(ix, c["ls"], c["le"], ["sum"])
    for i in range(len(c["ls"])):
        m = (c["rs"] < c["le"][i]) & (c["re"] > c["ls"][i])
        assert out[i, 0] == math.fsum(c["scores"][m])


def test_tile_modes_agree(orc, monkeypatch):
    """TMA-staged shared-memory windows vs plain global binary search: identical results."""
    import granges_b200 as gb

    c = random_case(11, 20000, 30000, span=400000, sorted_left=True, null_frac=0.1)
    outs = []
    for mode in ("1", "0"):
        monkeypatch.setenv("GRB_TILE_MODE", mode)
        ctx = gb.Context(0)
        ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
        outs.append((ctx.count_overlaps(ix, c["ls"], c["le"]), ctx.map_joins(ix, c["ls"], c["le"], ["mean", "min", "median"])))
    assert (outs[0][0] == outs[1][0]).all()
    assert (outs[0][1][0] == outs[1][1][0]).all() and (outs[0][1][1] == outs[1][1][1]).all()
    t = orc.Tree(c["rs"], c["re"])
    want = orc.map_joins(t, None, None, c["ls"], c["le"], ["count"])[0][:, 0]
    assert (outs[0][0] == want).all()


def test_custom_data_index(eng, orc):
    c = random_case(9, 500, 800, maxlen=5, span=300)
    idx = np.random.default_rng(1).permutation(800).astype(np.uint64)
    t = orc.Tree(c["rs"], c["re"], idx)
    ix = eng.build(c["rs"], c["re"], idx, c["scores"], None)
    off, gi, gs, ge = eng.left_overlaps(ix, c["ls"], c["le"])
    woff, wi, ws, we = orc.left_overlaps(t, c["ls"], c["le"])
    assert (off == woff).all() and (gi == wi).all() and (gs == ws).all() and (ge == we).all()
    out, ov = eng.map_joins(ix, c["ls"], c["le"], ["mean", "max"])
    want, wv = orc.map_joins(t, c["scores"], None, c["ls"], c["le"], ["mean", "max"])
    assert (ov == wv).all() and np.allclose(out, want, rtol=1e-12, atol=0)
    s, e, d = ix.sorted_ranges()
    o = np.lexsort((idx, c["re"], c["rs"]))
    assert (s == c["rs"][o]).all() and (e == c["re"][o]).all() and (d == idx[o]).all()


@pytest.mark.parametrize("kw", [dict(), dict(long_frac=0.02), dict(maxlen=4, span=200)], ids=["plain", "long", "ties"])
def test_pairs_vs_oracle(eng, orc, kw):
    c = random_case(31, 3000, 5000, **kw)
    t = orc.Tree(c["rs"], c["re"])
    ix = eng.build(c["rs"], c["re"])
    res = eng.ctx.left_overlaps(ix, c["ls"], c["le"], widths=True)
    off, gi, gs, ge, w = res
    woff, wi, ws, we = orc.left_overlaps(t, c["ls"], c["le"])
    assert (off == woff).all() and (gi == wi).all() and (gs == ws).all() and (ge == we).all()
    lid = np.repeat(np.arange(len(c["ls"])), np.diff(off.astype(np.int64)))
    ww = np.minimum(c["le"][lid], ge).astype(np.int64) - np.maximum(c["ls"][lid], gs).astype(np.int64)
    assert (w == ww).all() and (ww > 0).all()


def test_median_heavy_groups(eng, orc):
    """Groups with more members than the in-warp buffer (256) take the radix-select path."""
    rng = np.random.default_rng(4)
    nr = 6000
    rs = rng.integers(0, 2000, nr).astype(np.uint32)
    re = (rs + rng.integers(1, 800, nr)).astype(np.uint32)
    sc = np.round(rng.standard_normal(nr) * 50) / 8  # many ties, both signs, +-0
    sc[:7] = 0.0
    sc[7:14] = -0.0
    valid = (rng.random(nr) > 0.2).astype(np.uint8)
    ls = rng.integers(0, 2500, 300).astype(np.uint32)
    le = (ls + rng.integers(1, 1500, 300)).astype(np.uint32)
    c = dict(ls=ls, le=le, rs=rs, re=re, scores=sc, valid=valid)
    ix, t = run_case(eng, orc, c, ops=["median", "min", "max", "count", "mean"])
    cnt = eng.count(ix, ls, le)
    assert cnt.max() > 1000


def test_hg38_like_100k(eng, orc):
    """BASELINE.json configs[0]: 100k x 100k random ranges, lengths U{1..999} (chr1-sized seqname)."""
    rng = np.random.default_rng(13)
    n, span = 100_000, 5_000_000
    ln = rng.integers(1, 1000, n)
    rs = (rng.random(n) * (span - ln)).astype(np.uint32)
    re = (rs + ln).astype(np.uint32)
    ln = rng.integers(1, 1000, n)
    ls = np.sort((rng.random(n) * (span - ln)).astype(np.uint32))
    le = (ls + ln).astype(np.uint32)
    c = dict(ls=ls, le=le, rs=rs, re=re, scores=rng.random(n), valid=None)
    run_case(eng, orc, c, ops=["mean", "sum", "min", "max", "median", "count"])


def test_batch_multi_seqname(eng, orc):
    """Whole-GRanges call: 5 seqnames (one absent in the right, one without lefts) in one launch."""
    import granges_b200 as gb

    ctx = eng.ctx
    cases = [random_case(100 + s, nl, nr, span=50000, sorted_left=True) for s, (nl, nr) in
             enumerate([(3000, 2000), (0, 500), (1500, 0), (2048, 4000), (1, 1)])]
    idxs, trees = [], []
    for s, c in enumerate(cases):
        if s == 2:
            idxs.append(None)  # seqname absent in the right
            trees.append(None)
        else:
            idxs.append(ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"]))
            trees.append(orc.Tree(c["rs"], c["re"]))
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    ops = ["mean", "count", "max", "median", "sum"]
    out, ov = ctx.map_joins_batch(idxs, off, ls, le, ops)
    cnt = ctx.count_overlaps_batch(idxs, off, ls, le)
    keep, nk = ctx.filter_overlaps_batch(idxs, off, ls, le)
    akeep, ank = ctx.filter_overlaps_batch(idxs, off, ls, le, anti=True)
    for s, c in enumerate(cases):
        sl = slice(int(off[s]), int(off[s + 1]))
        want, wv = orc.map_joins(trees[s], c["scores"] if trees[s] is not None else None, None, c["ls"], c["le"], ops)
        assert (ov[sl] == wv).all()
        m = wv == 1
        assert np.allclose(out[sl][m], want[m], rtol=1e-12, atol=0)
        assert (out[sl][:, 2][wv[:, 2] == 1] == want[:, 2][wv[:, 2] == 1]).all()
        assert (cnt[sl] == want[:, 1]).all()
        assert (keep[sl] == (want[:, 1] > 0)).all() and (akeep[sl] == (want[:, 1] == 0)).all()
    assert nk == int(keep.sum()) and ank == int(akeep.sum()) and nk + ank == len(ls)


def test_device_resident_buffers(eng, orc):
    """torch CUDA tensors in, torch CUDA tensors out (no host staging)."""
    import torch

    c = random_case(55, 5000, 5000, span=60000, sorted_left=True)
    ix = eng.build(c["rs"], c["re"], None, c["scores"], None)
    dls = torch.from_numpy(c["ls"].astype(np.int32)).cuda()
    dle = torch.from_numpy(c["le"].astype(np.int32)).cuda()
    out, ov = eng.ctx.map_joins(ix, dls, dle, ["mean", "count"])
    eng.ctx.sync()
    assert out.is_cuda
    h, hv = eng.ctx.map_joins(ix, c["ls"], c["le"], ["mean", "count"])
    assert (out.cpu().numpy() == h).all() and (ov.cpu().numpy() == hv).all()


def test_windows_large_property(eng, orc):
    """hg38-sized window generation: device closed form == the reference's loop (oracle)."""
    lens = [248_956_422 // 50, 1_000_003, 999, 1000, 1001, 1]
    for width, step, chop in [(1000, 500, False), (1000, 500, True), (1000, None, False), (131123, 10001, True), (7, 3, False)]:
        got = eng.windows(lens, width, step, chop)
        want = orc.windows(lens, width, step, chop)
        for g, w in zip(got, want):
            assert len(g) == len(w) and (g == w).all()


def test_windows_then_map(eng, orc):
    """windows -> left_overlaps -> map_joins without leaving the device (granges windows | granges map)."""
    rng = np.random.default_rng(8)
    L, nr = 2_000_000, 50_000
    ln = rng.integers(1, 1000, nr)
    rs = (rng.random(nr) * (L - ln)).astype(np.uint32)
    re = (rs + ln).astype(np.uint32)
    sc = rng.random(nr)
    ix = eng.build(rs, re, None, sc, None)
    w = eng.ctx.windows_dev([L], 1000, 500, False)
    import ctypes as C

    import granges_b200._capi as capi

    ps, pe = w.dev_ptrs()
    n = len(w)
    out = np.empty((n, 1))
    ov = np.empty((n, 1), np.uint8)
    ops = (C.c_int * 1)(4)
    rc = capi.lib().grb_map_joins_f64(eng.ctx.h, ix.h, ps, pe, n, ops, 1, out.ctypes.data, ov.ctypes.data)
    assert rc == 0
    q, s, e = w.copy()
    want, wv = orc.map_joins(orc.Tree(rs, re), sc, None, s, e, ["mean"])
    assert (ov == wv).all() and np.allclose(out[wv == 1], want[wv == 1], rtol=1e-12, atol=0)


def test_errors(eng):
    import granges_b200 as gb

    ctx = eng.ctx
    with pytest.raises(gb.GRangesError) as ei:
        ctx.index_build([5], [5])
    assert ei.value.code == -2  # end <= start: the reference asserts (ranges/mod.rs:109)
    with pytest.raises(gb.GRangesError) as ei:
        ctx.index_build([0], [2**31 + 5])
    assert ei.value.code == -2  # does not fit i32 (ranges/coitrees.rs:53-54)
    ix = ctx.index_build([0, 10], [5, 20])
    with pytest.raises(gb.GRangesError) as ei:
        ctx.map_joins(ix, [0], [10], ["mean"])
    assert ei.value.code == -3  # NoDataContainer
    out, ov = ctx.map_joins(ix, [0], [10], ["count"])
    assert out[0, 0] == 1
    with pytest.raises(gb.GRangesError) as ei:
        ix.set_scores([1.0])  # data index 1 out of range
    assert ei.value.code == -7
    ix.set_scores([1.0, float("nan")])
    with pytest.raises(gb.GRangesError) as ei:
        ctx.map_joins(ix, [0], [30], ["median"])
    assert ei.value.code == -6  # the reference panics on NaN in median
    out, ov = ctx.map_joins(ix, [0], [30], ["sum", "min"])
    assert np.isnan(out[0, 0]) and out[0, 1] == 1.0
    with pytest.raises(gb.GRangesError):
        ctx.map_joins(ix, [0], [30], [99])


def test_unsorted_index_large_sort(eng, orc):
    """Index build from shuffled input exercises every radix pass (coordinates up to 2^31-1)."""
    rng = np.random.default_rng(21)
    n = 200_000
    rs = rng.integers(0, 2**31 - 2000, n).astype(np.uint32)
    re = (rs + rng.integers(1, 1000, n)).astype(np.uint32)
    ix = eng.build(rs, re)
    s, e, d = ix.sorted_ranges()
    o = np.lexsort((np.arange(n), re, rs))
    assert (s == rs[o]).all() and (e == re[o]).all() and (d == o).all()
    ls = np.sort(rng.integers(0, 2**31 - 5000, 50_000).astype(np.uint32))
    le = ls + 4000
    got = eng.count(ix, ls, le)
    es = np.sort(re)
    want = np.searchsorted(np.sort(rs), le, side="left") - np.searchsorted(es, ls, side="right")
    assert (got == want).all()


@pytest.mark.parametrize("op", ["mean", "sum", "sum-not-empty", "count", "count,mean,sum,sum-not-empty", "mean,mean"])
def test_pipeline_kernel_vs_tile_kernel(orc, monkeypatch, op):
    """Single-reduction maps run through the persistent TMA pipeline (k_map_pipe); it must agree bit
    for bit with the one-tile-per-CTA kernel and with the oracle -- ragged tiles at seqname
    boundaries, an absent seqname, tiles whose windows do not fit (sparse lefts over dense rights),
    unsorted lefts, and enough tiles that every ring slot is reused many times."""
    import granges_b200 as gb

    rng = np.random.default_rng(17)
    specs = [(40_000, 30_000, 900_000, True), (1, 5, 100, True), (2_500, 0, 1000, True), (5_000, 200_000, 3_000_000, True),
             (3_000, 3_000, 60_000, False), (1023, 2000, 30_000, True), (1025, 2000, 30_000, True)]
    cases = [random_case(300 + i, nl, nr, span=span, maxlen=999, sorted_left=srt) for i, (nl, nr, span, srt) in enumerate(specs)]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GRB_PIPE_MODE", mode)
        ctx = gb.Context(0)
        idxs = [None if i == 2 else ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"]) for i, c in enumerate(cases)]
        res[mode] = ctx.map_joins_batch(idxs, off, ls, le, op.split(","))
    assert (res["1"][1] == res["0"][1]).all()
    assert (res["1"][0][res["1"][1] == 1] == res["0"][0][res["0"][1] == 1]).all()
    for i, c in enumerate(cases):
        sl = slice(int(off[i]), int(off[i + 1]))
        t = None if i == 2 else orc.Tree(c["rs"], c["re"])
        want, wv = orc.map_joins(t, None if t is None else c["scores"], None, c["ls"], c["le"], op.split(","))
        assert (res["1"][1][sl] == wv).all()
        m = wv == 1
        assert np.allclose(res["1"][0][sl][m], want[m], rtol=1e-12, atol=0)


def test_sharded_units_bit_identical(eng, orc):
    """Multi-GPU units (a run of lefts + the rights its coordinate bounds admit) give exactly the bits of
    the unsharded call: sums are exact, everything else is order-free."""
    from granges_b200 import sharding

    c = random_case(71, 20_000, 30_000, span=2_000_000, maxlen=999, sorted_left=True)
    ops = ["mean", "sum", "min", "max", "median", "count"]
    ix = eng.build(c["rs"], c["re"], None, c["scores"], None)
    whole, wv = eng.map_joins(ix, c["ls"], c["le"], ops)
    off = np.array([0, len(c["ls"])], np.uint64)
    units = sharding.plan(off, [len(c["rs"])], 3, c["ls"], c["le"])
    assert len(units) == 3
    parts, vparts = [], []
    for u in units:
        a, b = u["left_begin"], u["left_end"]
        m = sharding.unit_right_mask(c["rs"], c["re"], u["right_lo"], u["right_hi"])
        assert m.sum() < len(m)
        sub = eng.build(c["rs"][m], c["re"][m], None, c["scores"][m], None)
        o, v = eng.map_joins(sub, c["ls"][a:b], c["le"][a:b], ops)
        parts.append(o)
        vparts.append(v)
    got = sharding.concat_in_order(units, parts)
    gv = sharding.concat_in_order(units, vparts)
    assert (gv == wv).all()
    assert (got[wv == 1] == whole[wv == 1]).all()


def test_invalid_left_ranges_are_reported(eng):
    """The reference cannot construct a range with end <= start (RangeEmpty::new asserts, ranges/mod.rs:30) and
    panics on coordinates beyond i32 (ranges/coitrees.rs:53-54): the kernels flag such lefts, answer them as empty
    groups (no out-of-range access) and the call reports GRB_ERR_INVALID_RANGE."""
    import granges_b200 as gb

    ctx = eng.ctx
    c = random_case(123, 3000, 4000, span=50_000, sorted_left=True)
    ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"])
    good, gv = ctx.map_joins(ix, c["ls"], c["le"], ["mean"])
    for kind in ("empty", "beyond-i32", "wrap"):
        ls, le = c["ls"].copy(), c["le"].copy()
        if kind == "empty":
            le[1500] = ls[1500]
        elif kind == "beyond-i32":
            le[1500] = np.uint32(2**31 + 7)
        else:  # coordinates that wrap when 1 is added: must not index outside any staged slice
            ls[1500] = np.uint32(2**32 - 1)
            le[1500] = np.uint32(2**32 - 1)
        for ops in (["mean"], ["min", "median", "sum"]):  # pipelined kernel / tile + sweep kernels
            with pytest.raises(gb.GRangesError) as ei:
                ctx.map_joins(ix, ls, le, ops)
            assert ei.value.code == -2
        with pytest.raises(gb.GRangesError):
            ctx.count_overlaps(ix, ls, le)
        with pytest.raises(gb.GRangesError):
            ctx.left_overlaps(ix, ls, le)
    # the flag is cleared by the failing call: the next valid call succeeds and is unchanged
    again, av = ctx.map_joins(ix, c["ls"], c["le"], ["mean"])
    assert (av == gv).all() and (again[gv == 1] == good[gv == 1]).all()


def test_index_build_batch(eng, orc):
    """grb_index_build_batch (copies of one seqname overlap the indexing of the previous) == per-seqname builds."""
    ctx = eng.ctx
    cases = [random_case(400 + s, 2000, nr, span=80_000, null_frac=nf) for s, (nr, nf) in enumerate([(5000, 0.0), (0, 0.0), (3000, 0.2), (1, 0.0)])]
    rights = [(c["rs"], c["re"], c["scores"], c["valid"]) for c in cases]
    idxs = ctx.index_build_batch(rights)
    assert [len(ix) for ix in idxs] == [len(c["rs"]) for c in cases]
    ops = ["mean", "median", "count", "sum"]
    for c, ix in zip(cases, idxs):
        single = eng.build(c["rs"], c["re"], None, c["scores"], c["valid"])
        a = ctx.map_joins(ix, c["ls"], c["le"], ops)
        b = ctx.map_joins(single, c["ls"], c["le"], ops)
        assert (a[1] == b[1]).all() and (a[0][a[1] == 1] == b[0][b[1] == 1]).all()
    # no data container at all
    bare = ctx.index_build_batch([(c["rs"], c["re"]) for c in cases])
    assert (ctx.count_overlaps(bare[0], cases[0]["ls"], cases[0]["le"]) == ctx.count_overlaps(idxs[0], cases[0]["ls"], cases[0]["le"])).all()


def test_host_buffers_chunked_pipeline(eng):
    """Large host-buffer calls run seqname chunks through an upload / compute / download pipeline; the result must be
    identical to the device-buffer call (single launch sequence), also with several reductions and ragged seqnames."""
    import torch

    ctx = eng.ctx
    rng = np.random.default_rng(5)
    nls = [900_001, 1_200_003, 0, 1_500_000, 700_002, 999_999]
    idxs, lss, les = [], [], []
    for s, nl in enumerate(nls):
        nr = 400_000 + 1000 * s
        rs = np.sort(rng.integers(0, 50_000_000, nr)).astype(np.uint32)
        re = (rs + rng.integers(1, 1000, nr)).astype(np.uint32)
        idxs.append(None if s == 4 else ctx.index_build(rs, re).set_scores(rng.random(nr)))
        ls = np.sort(rng.integers(0, 50_000_000, nl)).astype(np.uint32)
        lss.append(ls)
        les.append((ls + rng.integers(1, 1000, nl)).astype(np.uint32))
    off = np.concatenate([[0], np.cumsum(nls)]).astype(np.uint64)
    ls, le = np.concatenate(lss), np.concatenate(les)
    assert len(ls) >= 1 << 22
    dls, dle = torch.from_numpy(ls.view(np.int32)).cuda(), torch.from_numpy(le.view(np.int32)).cuda()
    for ops in (["mean"], ["max", "mean", "count"], ["overlap-bp", "weighted-mean", "median"]):
        host_out, host_v = ctx.map_joins_batch(idxs, off, ls, le, ops)            # chunked host path
        dev_out, dev_v = ctx.map_joins_batch(idxs, off, dls, dle, ops)            # device path
        ctx.sync()
        assert (host_v == dev_v.cpu().numpy()).all()
        m = host_v == 1
        assert (host_out[m] == dev_out.cpu().numpy()[m]).all()


def test_filter_select_compaction(eng, orc):
    """filter_overlaps as an order-preserving compaction: kept indices ascending, equal to the flagged rows."""
    ctx = eng.ctx
    cases = [random_case(600 + s, nl, nr, span=30_000) for s, (nl, nr) in enumerate([(5000, 300), (0, 10), (3000, 0), (4097, 2000)])]
    idxs = [None if len(c["rs"]) == 0 else ctx.index_build(c["rs"], c["re"]) for c in cases]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    for anti in (False, True):
        keep, nk = ctx.filter_overlaps_batch(idxs, off, ls, le, anti=anti)
        sel = ctx.filter_overlaps_select(idxs, off, ls, le, anti=anti)
        assert len(sel) == nk and (sel == np.flatnonzero(keep)).all()
    want = np.concatenate([orc.filter_overlaps(None if ix is None else orc.Tree(c["rs"], c["re"]), c["ls"], c["le"]) for c, ix in zip(cases, idxs)])
    assert (ctx.filter_overlaps_select(idxs, off, ls, le) == np.flatnonzero(want)).all()


def test_many_seqnames(eng, orc):
    """More seqnames than the pipelined kernel keeps descriptors for (32): several launches, same answers."""
    ctx = eng.ctx
    rng = np.random.default_rng(77)
    nseq = 75
    cases = [random_case(900 + s, int(rng.integers(0, 3000)), int(rng.integers(0, 2000)), span=40_000, sorted_left=True) for s in range(nseq)]
    idxs = [None if len(c["rs"]) == 0 else ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"]) for c in cases]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    out, ov = ctx.map_joins_batch(idxs, off, ls, le, ["mean"])
    for s, c in enumerate(cases):
        sl = slice(int(off[s]), int(off[s + 1]))
        t = None if idxs[s] is None else orc.Tree(c["rs"], c["re"])
        want, wv = orc.map_joins(t, None if t is None else c["scores"], None, c["ls"], c["le"], ["mean"])
        assert (ov[sl] == wv).all()
        m = wv[:, 0] == 1
        assert np.allclose(out[sl][m], want[m], rtol=1e-12, atol=0)


@pytest.mark.parametrize("kind", ["compact", "wide"])
@pytest.mark.parametrize("op", ["mean", "sum"])
def test_large_groups_exact_128bit_path(eng, orc, op, kind):
    """Groups of more than 2^(63-W) members (1024 for 53-bit scores) leave the 64-bit wrapped-prefix fast path and
    rebuild the exact 128-bit prefixes from the per-256 bases -- in the pipelined kernel (k = 1) and the tile kernel."""
    import math

    rng = np.random.default_rng(41)
    nr = 20_000
    rs = np.sort(rng.integers(0, 3000, nr)).astype(np.uint32)
    re = (rs + rng.integers(1, 900, nr)).astype(np.uint32)
    # compact: multiples of 2^-53 below 1 (W = 53, the pipelined kernel's prefixes); wide: 53-bit mantissas at many
    # exponents and both signs (128-bit prefixes per right, tile kernel)
    sc = rng.random(nr) if kind == "compact" else rng.random(nr) * 1000.0 - 300.0
    ls = np.sort(rng.integers(0, 3500, 4000)).astype(np.uint32)
    le = (ls + rng.integers(1, 2000, 4000)).astype(np.uint32)
    ix = eng.build(rs, re, None, sc, None)
    assert ix.exact_sums
    cnt = eng.count(ix, ls, le)
    assert cnt.max() > 5000 and (cnt > 1024).mean() > 0.5
    out, ov = eng.map_joins(ix, ls, le, [op])                 # pipelined kernel
    out2, ov2 = eng.map_joins(ix, ls, le, [op, "count"])      # tile kernel (two reductions)
    assert (ov == ov2[:, :1]).all() and (out[ov == 1] == out2[:, :1][ov == 1]).all()
    for i in range(0, len(ls), 37):
        m = (rs < le[i]) & (re > ls[i])
        exact = math.fsum(sc[m])
        want = exact if op == "sum" else (exact / int(m.sum()) if m.any() else None)
        if want is None:
            assert ov[i, 0] == 0
        else:
            assert ov[i, 0] == 1 and out[i, 0] == want, (i, out[i, 0], want)


RANK_SWEEP_CASES = {
    "sorted": dict(nl=6000, nr=9000, span=120000, maxlen=999, sorted_left=True),
    "unsorted": dict(nl=3000, nr=5000, span=60000, maxlen=999),
    "nulls-ties": dict(nl=5000, nr=7000, span=90000, maxlen=500, sorted_left=True, null_frac=0.4, score_kind="int"),
    "long-rights": dict(nl=4000, nr=20000, span=300000, maxlen=300, long_frac=0.01, sorted_left=True),
    "mid-dense": dict(nl=4000, nr=12000, span=60000, maxlen=500, sorted_left=True),
    "dense": dict(nl=2500, nr=40000, span=20000, maxlen=400, sorted_left=True),
    "very-dense": dict(nl=1200, nr=60000, span=6000, maxlen=900, sorted_left=True, null_frac=0.1),
    "ultra-dense": dict(nl=300, nr=100000, span=4000, maxlen=1500, sorted_left=True, null_frac=0.05, score_kind="int"),
    "mixed-sign": dict(nl=3000, nr=3000, span=30000, maxlen=700, sorted_left=True, score_kind="mixed"),
}


@pytest.mark.parametrize("name", list(RANK_SWEEP_CASES))
def test_rank_sweep_modes(orc, monkeypatch, name):
    """MIN / MAX / MEDIAN through the staged rank sweep (k_sweep3: slices of ends / ranks in shared memory,
    halved runs, global fallback), the same kernel without staging, and the older score sweep (k_sweep2):
    bit-identical to each other and to the oracle -- slices that do not fit, groups beyond the in-group
    buffers (radix-select path), None scores, ties, unsorted lefts."""
    import granges_b200 as gb

    kw = dict(RANK_SWEEP_CASES[name])
    nl, nr = kw.pop("nl"), kw.pop("nr")
    c = random_case(4242 + nl, nl, nr, **kw)
    ops = ["min", "count", "median", "mean", "max"]
    res = {}
    for mode in ("1", "2", "0"):
        monkeypatch.setenv("GRB_SWEEP_MODE", mode)
        ctx = gb.Context(0)
        ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
        res[mode] = ctx.map_joins(ix, c["ls"], c["le"], ops)
        # a second call reuses the member tables
        again = ctx.map_joins(ix, c["ls"], c["le"], ["median"])
        assert (again[1][:, 0] == res[mode][1][:, 2]).all()
        m = again[1][:, 0] == 1
        assert (again[0][m, 0] == res[mode][0][m, 2]).all()
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    for mode in res:
        out, ov = res[mode]
        assert (ov == wv).all(), mode
        for col, op in enumerate(ops):
            m = wv[:, col] == 1
            if op == "mean":
                assert np.allclose(out[m, col], want[m, col], rtol=1e-12, atol=0), (mode, op)
            else:
                assert (out[m, col] == want[m, col]).all(), (mode, op)


def test_rank_sweep_batch_with_absent_seqnames(eng, orc):
    """Whole-GRanges MIN / MAX / MEDIAN call: an absent seqname, an empty index and one without lefts."""
    cases = [random_case(900 + i, nl, nr, span=50000, maxlen=800, sorted_left=True, null_frac=nf)
             for i, (nl, nr, nf) in enumerate([(3000, 4000, 0.0), (1500, 0, 0.0), (0, 500, 0.0), (2100, 2600, 0.3), (700, 900, 0.0)])]
    ops = ["median", "max", "min"]
    idxs = [None if i == 1 else eng.build(c["rs"], c["re"], None, c["scores"], c["valid"]) for i, c in enumerate(cases)]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    out, ov = eng.ctx.map_joins_batch(idxs, off, ls, le, ops)
    for i, c in enumerate(cases):
        sl = slice(int(off[i]), int(off[i + 1]))
        if i == 1:
            assert (ov[sl] == 0).all()
            continue
        t = orc.Tree(c["rs"], c["re"])
        want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
        assert (ov[sl] == wv).all()
        m = wv == 1
        assert (out[sl][m] == want[m]).all()


@pytest.mark.parametrize("name", ["sorted", "unsorted", "long-rights", "mid-dense", "nulls-ties"])
@pytest.mark.parametrize("ops", [["overlap-bp"], ["min", "overlap-bp", "mean", "median"], ["overlap-bp", "weighted-mean"]],
                         ids=["bp", "bp+members", "bp+wm"])
def test_overlap_bp_closed_form(eng, orc, name, ops):
    """OVERLAP_BP from the prefix sums of the sorted starts / ends (k_overlap_bp, no member visited) is exact,
    alone, beside member reductions and beside a weighted mean."""
    kw = dict(RANK_SWEEP_CASES[name])
    nl, nr = kw.pop("nl"), kw.pop("nr")
    c = random_case(777 + nr, nl, nr, **kw)
    ix = eng.build(c["rs"], c["re"], None, c["scores"], c["valid"])
    out, ov = eng.map_joins(ix, c["ls"], c["le"], ops)
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    assert (ov == wv).all()
    for col, op in enumerate(ops):
        m = wv[:, col] == 1
        if op in ("mean", "weighted-mean"):
            assert np.allclose(out[m, col], want[m, col], rtol=1e-12, atol=0), op
        else:
            assert (out[m, col] == want[m, col]).all(), op
    # brute force for the bp column
    col = ops.index("overlap-bp")
    for i in range(0, nl, max(1, nl // 200)):
        w = np.minimum(c["re"], c["le"][i]).astype(np.int64) - np.maximum(c["rs"], c["ls"][i]).astype(np.int64)
        assert out[i, col] == w[w > 0].sum()


@pytest.mark.parametrize("kind", ["plain", "nulls-int", "long-mixed", "dense-ties", "wide", "sorted-left", "nulls-sorted"])
def test_weighted_mean_closed_form(orc, monkeypatch, kind):
    """WEIGHTED_MEAN (examples/weighted_mean.rs:4-31) from exact prefixes of score * coordinate (k_weighted_mean) against
    the member sweep (GRB_SWEEP_MODE=0) and the oracle; score ranges too wide for the fixed point ("wide") fall back
    to the sweep on their own."""
    import granges_b200 as gb

    c = random_case(4100, 4000, 6000, **dict(KINDS[kind], span=KINDS[kind].get("span", 40000)))
    ops = ["weighted-mean", "overlap-bp", "count", "mean"]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GRB_SWEEP_MODE", mode)
        ctx = gb.Context(0)
        ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
        res[mode] = ctx.map_joins(ix, c["ls"], c["le"], ops)
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    amax = np.abs(c["scores"]).max() if len(c["scores"]) else 0.0
    for mode in res:
        out, ov = res[mode]
        assert (ov == wv).all()
        assert (out[:, 1] == want[:, 1]).all() and (out[:, 2] == want[:, 2]).all()
        # |weighted mean| <= max |score|; both sides round a sum of products: 1e-12 of that scale
        assert (np.abs(out[:, 0] - want[:, 0]) <= 1e-12 * amax).all(), mode
    # brute force of the closed form's exactness on a few lefts: sum(score * width) / sum(width) in exact arithmetic
    from fractions import Fraction

    valid = np.ones(len(c["rs"]), bool) if c["valid"] is None else c["valid"].astype(bool)
    if kind not in ("wide", "long-mixed"):  # those score ranges leave no room for score * coordinate: swept, not exact
        for i in range(0, len(c["ls"]), 400):
            w = np.minimum(c["re"], c["le"][i]).astype(np.int64) - np.maximum(c["rs"], c["ls"][i]).astype(np.int64)
            m = (w > 0) & valid
            if not m.any():
                assert res["1"][0][i, 0] == 0.0
                continue
            num = sum(Fraction(float(s)) * int(x) for s, x in zip(c["scores"][m], w[m]))
            exact = float(num) / float(int(w[m].sum()))
            assert res["1"][0][i, 0] == exact


@pytest.mark.parametrize("ops", [["mean"], ["min", "mean", "median", "count"], ["weighted-mean", "overlap-bp"]], ids=["mean", "mixed", "widths"])
def test_map_whole_equals_two_calls(eng, orc, ops):
    """grb_map_joins_whole_f64 (`granges map` in one call, pipelined per seqname) returns exactly what
    grb_index_build_batch + grb_map_joins_f64_batch return -- seqnames without rights, without lefts, None scores."""
    specs = [(60_000, 50_000, 3_000_000, 0.0), (1500, 0, 10_000, 0.0), (0, 500_000, 10_000_000, 0.0), (45_000, 70_000, 5_000_000, 0.2),
             (700, 900, 40_000, 0.0), (30_000, 20_000, 900_000, 0.0), (5, 3, 100, 0.0), (20_000, 25_000, 1_000_000, 0.0),
             (20_000, 25_000, 1_000_000, 0.0), (9_000, 12_000, 400_000, 0.5)]
    cases = [random_case(1200 + i, nl, nr, span=span, maxlen=900, sorted_left=True, null_frac=nf) for i, (nl, nr, span, nf) in enumerate(specs)]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    rights = [(c["rs"], c["re"], c["scores"], c["valid"] if c["valid"] is not None else np.ones(len(c["rs"]), np.uint8)) for c in cases]
    got, gv = eng.ctx.map_whole(rights, off, ls, le, ops)
    idxs = eng.ctx.index_build_batch(rights)
    want, wv = eng.ctx.map_joins_batch(idxs, off, ls, le, ops)
    assert (gv == wv).all()
    m = wv == 1
    assert (got[m] == want[m]).all()
    # and the oracle on one seqname
    c = cases[3]
    sl = slice(int(off[3]), int(off[4]))
    t = orc.Tree(c["rs"], c["re"])
    o_want, o_wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    assert (gv[sl] == o_wv).all()
    for col, op in enumerate(ops):
        mm = o_wv[:, col] == 1
        if op in ("mean", "weighted-mean"):
            assert np.allclose(got[sl][mm, col], o_want[mm, col], rtol=1e-12, atol=0)
        else:
            assert (got[sl][mm, col] == o_want[mm, col]).all()


def test_map_whole_mgpu_and_pageable_staging(eng, orc):
    """grb_map_joins_whole_f64_mgpu (one process, one host thread + context per GPU, contiguous runs of seqnames, every
    GPU writing its rows of the caller's arrays) returns the single-GPU call's bits on every GPU count the box offers;
    so do pageable host buffers large enough for the page-locked staging rings and page-locked ones."""
    import torch

    import granges_b200 as gb

    specs = [(400_000, 300_000, 30_000_000, 0.0), (0, 1000, 10_000, 0.0), (350_000, 500_000, 40_000_000, 0.1), (1200, 0, 10_000, 0.0),
             (300_000, 280_000, 20_000_000, 0.0), (5, 3, 100, 0.0), (280_000, 300_000, 25_000_000, 0.0)]
    cases = [random_case(1300 + i, nl, nr, span=span, maxlen=900, sorted_left=True, null_frac=nf) for i, (nl, nr, span, nf) in enumerate(specs)]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    rights = [(c["rs"], c["re"], c["scores"], c["valid"] if c["valid"] is not None else np.ones(len(c["rs"]), np.uint8)) for c in cases]
    ops = ["mean", "count"]
    want, wv = eng.ctx.map_whole(rights, off, ls, le, ops)  # pageable numpy arrays: staged through the rings
    # the oracle on one seqname
    c = cases[2]
    sl = slice(int(off[2]), int(off[3]))
    o_want, o_wv = orc.map_joins(orc.Tree(c["rs"], c["re"]), c["scores"], c["valid"], c["ls"], c["le"], ops, nthreads=4)
    assert (wv[sl] == o_wv).all() and (want[sl][:, 1] == o_want[:, 1]).all()
    m = o_wv[:, 0] == 1
    assert np.allclose(want[sl][m, 0], o_want[m, 0], rtol=1e-12, atol=0)
    # page-locked buffers take the plain asynchronous copies
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    p_rights = [tuple(pin(x) for x in r) for r in rights]
    out_p, ov_p = pin(np.zeros_like(want)), pin(np.zeros_like(wv))
    eng.ctx.map_whole(p_rights, off, pin(ls), pin(le), ops, out=out_p, out_valid=ov_p)
    assert (ov_p == wv).all() and (out_p[wv == 1] == want[wv == 1]).all()
    # every GPU count
    ndev = torch.cuda.device_count()
    for n in sorted({1, min(2, ndev), ndev}):
        got, gv = gb.map_whole_mgpu(list(range(n)), rights, off, ls, le, ops)
        assert (gv == wv).all() and (got[wv == 1] == want[wv == 1]).all(), n
    got, gv = gb.map_whole_mgpu([0], p_rights, off, pin(ls), pin(le), ops, out=out_p, out_valid=ov_p)
    assert (gv == wv).all() and (got[wv == 1] == want[wv == 1]).all()
    with pytest.raises(gb.GRangesError):
        gb.map_whole_mgpu([ndev + 3], rights, off, ls, le, ops)


@pytest.mark.parametrize("name", list(RANK_SWEEP_CASES))
def test_minmax_without_a_sweep(orc, monkeypatch, name):
    """MIN / MAX with no median beside them come from the stabbing tables + block sparse table (k_minmax): the same
    bits as the member sweep (GRB_MINMAX_MODE=0) and the oracle -- long rights, dense groups spanning many blocks,
    None scores, ties, unsorted lefts, length-1 rights."""
    import granges_b200 as gb

    kw = dict(RANK_SWEEP_CASES[name])
    nl, nr = kw.pop("nl"), kw.pop("nr")
    c = random_case(5151 + nr, nl, nr, **kw)
    # sprinkle length-1 rights (their open and close events share a coordinate) and exact book-ends
    c["re"][::17] = c["rs"][::17] + 1
    ops = ["max", "mean", "min", "count"]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GRB_MINMAX_MODE", mode)
        ctx = gb.Context(0)
        ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
        n0 = ctx.launch_count
        res[mode] = ctx.map_joins(ix, c["ls"], c["le"], ops)
        n1 = ctx.launch_count
        res[mode + "b"] = ctx.map_joins(ix, c["ls"], c["le"], ["min"])  # tables are reused: far fewer launches than the first call
        assert ctx.launch_count - n1 < n1 - n0
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    abs_sums = orc.map_joins(t, np.abs(c["scores"]), c["valid"], c["ls"], c["le"], ["sum"])[0][:, 0]
    for mode in ("1", "0"):
        out, ov = res[mode]
        assert_map_equal(out, ov, want, wv, ops, abs_sums)
        ob, vb = res[mode + "b"]
        assert (vb[:, 0] == wv[:, 2]).all()
        m = wv[:, 2] == 1
        assert (ob[m, 0] == want[m, 2]).all()


def test_minmax_batch_with_absent_seqnames(eng, orc):
    cases = [random_case(950 + i, nl, nr, span=50000, maxlen=800, sorted_left=True, null_frac=nf)
             for i, (nl, nr, nf) in enumerate([(3000, 4000, 0.0), (1500, 0, 0.0), (0, 500, 0.0), (2100, 2600, 0.3), (700, 900, 0.0)])]
    ops = ["max", "min"]
    idxs = [None if i == 1 else eng.build(c["rs"], c["re"], None, c["scores"], c["valid"]) for i, c in enumerate(cases)]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    out, ov = eng.ctx.map_joins_batch(idxs, off, ls, le, ops)
    for i, c in enumerate(cases):
        sl = slice(int(off[i]), int(off[i + 1]))
        if i == 1:
            assert (ov[sl] == 0).all()
            continue
        t = orc.Tree(c["rs"], c["re"])
        want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
        assert (ov[sl] == wv).all()
        m = wv == 1
        assert (out[sl][m] == want[m]).all()


def test_closed_forms_at_the_coordinate_limits(eng, orc):
    """Coordinates up to 2^31 - 1 (the reference's i32 limit), lefts entirely before / after every right, single-range
    inputs: MIN / MAX (k_minmax), MEDIAN, OVERLAP_BP, WEIGHTED_MEAN, SUM, COUNT against the oracle."""
    top = 2**31 - 1
    rs = np.array([0, 5, top - 900, top - 500, top - 2, 1_000_000_000, 1_000_000_000, 1_000_000_010], np.uint32)
    re = np.array([top, 6, top - 100, top, top, 1_000_000_500, 1_000_000_001, top], np.uint32)
    sc = np.array([0.5, -2.25, 3.0, 1e6, -1e-3, 7.0, 7.0, 0.125])
    ls = np.array([0, 0, 6, top - 1, top - 600, 999_999_999, 1_000_000_000, 1_000_000_001, 2, top - 1000], np.uint32)
    le = np.array([1, top, 7, top, top - 400, 1_000_000_001, 1_000_000_001, 1_000_000_002, 5, top], np.uint32)
    ops = ["min", "max", "overlap-bp", "weighted-mean", "sum", "count"]
    for valid in (None, np.array([1, 1, 0, 1, 1, 0, 1, 1], np.uint8)):
        ix = eng.build(rs, re, None, sc, valid)
        t = orc.Tree(rs, re)
        want, wv = orc.map_joins(t, sc, valid, ls, le, ops)
        out, ov = eng.map_joins(ix, ls, le, ops)
        assert (ov == wv).all()
        for col, op in enumerate(ops):
            m = wv[:, col] == 1
            if op in ("weighted-mean", "sum"):
                assert np.allclose(out[m, col], want[m, col], rtol=1e-12, atol=0), op
            else:
                assert (out[m, col] == want[m, col]).all(), op
        want, wv = orc.map_joins(t, sc, valid, ls, le, ["median", "min"])
        out, ov = eng.map_joins(ix, ls, le, ["median", "min"])
        assert (ov == wv).all() and (out[wv == 1] == want[wv == 1]).all()
    # one right, lefts on every side of it
    ix = eng.build([100], [200], None, [4.5], None)
    out, ov = eng.map_joins(ix, [0, 100, 150, 199, 200, 50], [100, 101, 160, 300, 300, 250], ["min", "overlap-bp", "weighted-mean", "median"])
    assert ov.tolist() == [[0, 1, 1, 0], [1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1], [0, 1, 1, 0], [1, 1, 1, 1]]
    assert out[:, 1].tolist() == [0, 1, 10, 1, 0, 100]
    assert out[[1, 2, 3, 5], 0].tolist() == [4.5] * 4 and out[[1, 2, 3, 5], 2].tolist() == [4.5] * 4


def test_map_whole_degenerate_shapes(eng, orc):
    """One seqname; a genome where no seqname has lefts; rights without scores for a count."""
    c = random_case(31, 5000, 4000, span=200_000, maxlen=700, sorted_left=True)
    got, gv = eng.ctx.map_whole([(c["rs"], c["re"], c["scores"])], [0, len(c["ls"])], c["ls"], c["le"], ["mean", "max"])
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], None, c["ls"], c["le"], ["mean", "max"])
    assert (gv == wv).all()
    assert np.allclose(got[:, 0][wv[:, 0] == 1], want[:, 0][wv[:, 0] == 1], rtol=1e-12, atol=0)
    assert (got[:, 1][wv[:, 1] == 1] == want[:, 1][wv[:, 1] == 1]).all()
    out, ov = eng.ctx.map_whole([(c["rs"], c["re"], c["scores"])] * 3, [0, 0, 0, 0], np.zeros(0, np.uint32), np.zeros(0, np.uint32), ["mean"])
    assert out.shape == (0, 1)
    got, gv = eng.ctx.map_whole([(c["rs"], c["re"])], [0, len(c["ls"])], c["ls"], c["le"], ["count"])
    assert (got[:, 0] == orc.map_joins(t, None, None, c["ls"], c["le"], ["count"])[0][:, 0]).all()


@pytest.mark.parametrize("variant", ["sums", "no-sums-with-nulls", "median"])
def test_counts_through_the_pipeline(orc, monkeypatch, variant):
    """The (ub, pp, lb) pass in front of the member sweep / closed forms runs in the persistent pipeline when there are
    many tiles (k_map_pipe<UBLB>): bit-identical to the one-tile-per-CTA kernel (GRB_PIPE_MODE=0) and to the oracle --
    ragged seqname boundaries, an absent seqname, None scores (counts only), crowded bins, unsorted lefts."""
    import granges_b200 as gb

    nf = 0.3 if variant == "no-sums-with-nulls" else 0.0
    specs = [(330_000, 200_000, 9_000_000, True), (1, 5, 100, True), (2_500, 0, 1000, True), (250_000, 400_000, 6_000_000, True),
             (80_000, 3_000, 60_000_000, True), (30_000, 30_000, 600_000, False), (1025, 5000, 60, True)]
    cases = [random_case(800 + i, nl, nr, span=span, maxlen=999, sorted_left=srt, null_frac=nf) for i, (nl, nr, span, srt) in enumerate(specs)]
    off = np.concatenate([[0], np.cumsum([len(c["ls"]) for c in cases])]).astype(np.uint64)
    assert off[-1] >= 620_000
    ls = np.concatenate([c["ls"] for c in cases])
    le = np.concatenate([c["le"] for c in cases])
    ops = {"sums": ["min", "count", "mean", "max", "sum", "weighted-mean"], "no-sums-with-nulls": ["max", "overlap-bp", "count", "min"],
           "median": ["median", "mean", "overlap-bp"]}[variant]
    res = {}
    monkeypatch.setenv("GRB_UBLB_PIPE", "1")  # also where the default keeps the tile kernel (no sums)
    for mode in ("1", "0"):
        monkeypatch.setenv("GRB_PIPE_MODE", mode)
        ctx = gb.Context(0)
        idxs = [None if i == 2 else ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"]) for i, c in enumerate(cases)]
        res[mode] = ctx.map_joins_batch(idxs, off, ls, le, ops)
    assert (res["1"][1] == res["0"][1]).all()
    m = res["1"][1] == 1
    assert (res["1"][0][m] == res["0"][0][m]).all()
    for i in (0, 3, 5, 6):
        c = cases[i]
        sl = slice(int(off[i]), int(off[i + 1]))
        pick = np.arange(0, len(c["ls"]), 37)
        t = orc.Tree(c["rs"], c["re"])
        want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"][pick], c["le"][pick], ops)
        got, gv = res["1"][0][sl][pick], res["1"][1][sl][pick]
        assert (gv == wv).all()
        for col, op in enumerate(ops):
            mm = wv[:, col] == 1
            if op in ("mean", "sum", "weighted-mean"):
                assert np.allclose(got[mm, col], want[mm, col], rtol=1e-12, atol=0), op
            else:
                assert (got[mm, col] == want[mm, col]).all(), op
