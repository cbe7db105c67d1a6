"""The pipelined prefix-sum kernel (csrc/pipe.cu) on every input class the reference accepts at one speed
(src/commands.rs:524-542: `None` scores are dropped; src/data/operations.rs:60-70,87-95: plain f64 folds, so
decimal fractions, mixed signs, +-inf and NaN all go through): None scores (the NULLS variant), scores that need
96-bit fixed point (WIDE), non-finite scores (prefix counters + the IEEE rule), sorted lefts (staged windows) and
shuffled lefts (per-left global searches), single reductions and several at once, validity as bytes and as bits.
Every result is checked against the CPU oracle; variants that must agree bit for bit are compared directly."""
import numpy as np
import pytest

from tests.randcases import random_case
from tests.mapchecks import assert_map_equal, finite_abs, same_bits

pytestmark = pytest.mark.gpu

PIPE_OPS = ["mean", "sum", "sum-not-empty", "count"]


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g

    g.build()
    import granges_b200 as gb

    return gb.Context(0)


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


def oracle_refs(orc, c, ops):
    t = orc.Tree(c["rs"], c["re"])
    want, wv = orc.map_joins(t, c["scores"], c["valid"], c["ls"], c["le"], ops)
    abs_sums = orc.map_joins(t, finite_abs(c["scores"]), c["valid"], c["ls"], c["le"], ["sum"])[0][:, 0]
    nvalid = orc.map_joins(t, np.ones(len(c["rs"])), c["valid"], c["ls"], c["le"], ["sum"])[0][:, 0]
    return want, wv, abs_sums, nvalid


def check_all_shapes(ctx, orc, c):
    """Each reduction alone (the FAST kernels), all of them at once (PIPE_MULTI), bytes and bits."""
    ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
    want, wv, abs_sums, nvalid = oracle_refs(orc, c, PIPE_OPS)
    out, ov = ctx.map_joins(ix, c["ls"], c["le"], PIPE_OPS)
    assert_map_equal(out, ov, want, wv, PIPE_OPS, abs_sums, nvalid)
    nl = len(c["ls"])
    off = np.array([0, nl], np.uint64)
    for col, op in enumerate(PIPE_OPS):
        o1, v1 = ctx.map_joins(ix, c["ls"], c["le"], [op])
        assert_map_equal(o1, v1, want[:, col:col + 1], wv[:, col:col + 1], [op], abs_sums, nvalid)
        # single and multi column calls run the same arithmetic: identical bits
        m = wv[:, col] == 1
        assert same_bits(o1[m, 0], out[m, col])
        # validity as bits
        o2, bits = ctx.map_joins_batch([ix], off, c["ls"], c["le"], [op], valid_bits=True)
        got = np.unpackbits(bits.view(np.uint8), bitorder="little")[:nl]
        assert (got == wv[:, col]).all(), op
        assert same_bits(o2[m, 0], o1[m, 0])
    return ix


VARIANTS = {
    "plain": dict(),
    "nulls": dict(null_frac=0.3),
    "nulls-sparse": dict(null_frac=0.01),
    "decimal": dict(score_kind="decimal"),
    "decimal-nulls": dict(score_kind="decimal", null_frac=0.2),
    "signed-decimal": dict(score_kind="signed-decimal"),
    "wide96": dict(score_kind="wide96"),
    "nonfinite": dict(score_kind="nonfinite"),
    "nonfinite-nulls": dict(score_kind="nonfinite", null_frac=0.2),
    "nan": dict(score_kind="nan"),
    "int": dict(score_kind="int"),
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("shape", ["staged", "shuffled", "tiny", "dense"])
def test_pipe_variant_vs_oracle(ctx, orc, variant, shape):
    kw = dict(VARIANTS[variant])
    if shape == "staged":    # sorted lefts over a span that keeps every tile's windows inside the ring slots
        c = random_case(501, 6000, 9000, span=120_000, sorted_left=True, **kw)
    elif shape == "shuffled":  # unsorted lefts: the tiles' windows do not fit, every left searches global memory
        c = random_case(502, 5000, 20000, span=400_000, **kw)
    elif shape == "tiny":
        c = random_case(503, 37, 21, span=300, **kw)
    else:                    # crowded bins, long scans, groups of thousands (beyond the 64-bit fast path for 53-bit scores)
        c = random_case(504, 3000, 6000, span=60, maxlen=8, sorted_left=True, **kw)
    check_all_shapes(ctx, orc, c)


def test_sum_modes_agree_bit_for_bit(ctx, orc, monkeypatch):
    """compact (64-bit planes), wide96 (64 + 32-bit planes) and wide (i128 per right) are all the correctly rounded exact
    sum: forcing the wider formats on compact data must not change a bit, in the pipeline and in the tile kernel."""
    import granges_b200 as gb

    c = random_case(78, 5000, 8000, span=100_000, sorted_left=True, null_frac=0.1)
    res = {}
    for mode in ("1", "3", "2"):
        monkeypatch.setenv("GRB_SUM_MODE", mode)
        ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
        res[mode] = [ctx.map_joins(ix, c["ls"], c["le"], [op]) for op in ("mean", "sum")] + [ctx.map_joins(ix, c["ls"], c["le"], PIPE_OPS)]
    monkeypatch.delenv("GRB_SUM_MODE")
    for mode in ("3", "2"):
        for (a, av), (b, bv) in zip(res["1"], res[mode]):
            assert (av == bv).all() and same_bits(a[av.astype(bool)], b[bv.astype(bool)]), mode
    # and the tile kernel (pipeline off) agrees with the pipeline
    monkeypatch.setenv("GRB_PIPE_MODE", "0")
    ctx2 = gb.Context(0)
    for kw in (dict(), dict(score_kind="decimal", null_frac=0.2), dict(score_kind="nonfinite")):
        c2 = random_case(79, 4000, 6000, span=80_000, sorted_left=True, **kw)
        a, av = ctx.map_joins(ctx.index_build(c2["rs"], c2["re"]).set_scores(c2["scores"], c2["valid"]), c2["ls"], c2["le"], ["mean"])
        b, bv = ctx2.map_joins(ctx2.index_build(c2["rs"], c2["re"]).set_scores(c2["scores"], c2["valid"]), c2["ls"], c2["le"], ["mean"])
        assert (av == bv).all()
        m = av[:, 0] == 1
        assert same_bits(a[m], b[m])


def test_exact_sums_mixed_signs_are_correctly_rounded(ctx):
    """Under cancellation the prefix path returns the correctly rounded EXACT sum (math.fsum) -- where the reference's
    left-to-right f64 sum loses low bits (1e16 + 1 - 1e16 gives 0 there, 1 here): a documented deviation that stays
    inside 1e-12 * sum|x|."""
    import math

    rs = np.array([0, 1, 2, 50], np.uint32)
    re = np.array([10, 11, 12, 60], np.uint32)
    sc = np.array([1e16, 1.0, -1e16, 0.25])
    ix = ctx.index_build(rs, re).set_scores(sc)
    assert ix.exact_sums
    out, ov = ctx.map_joins(ix, np.array([0, 40], np.uint32), np.array([20, 70], np.uint32), ["sum"])
    assert out[0, 0] == 1.0 and out[1, 0] == 0.25
    c = random_case(6, 500, 800, span=4000, score_kind="signed-decimal", sorted_left=True)
    ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"])
    out, ov = ctx.map_joins(ix, c["ls"], c["le"], ["sum"])
    for i in range(len(c["ls"])):
        m = (c["rs"] < c["le"][i]) & (c["re"] > c["ls"][i])
        assert out[i, 0] == math.fsum(c["scores"][m])


def test_nulls_group_beyond_16_bit_counts(ctx, orc):
    """Groups of more than 65535 members: the staged low halves of the valid-score counts are not enough, the
    kernel reads the full counters."""
    rng = np.random.default_rng(9)
    nr = 70_000
    rs = rng.integers(0, 40, nr).astype(np.uint32)
    re = (rs + rng.integers(1, 30, nr)).astype(np.uint32)
    sc = rng.random(nr)
    valid = (rng.random(nr) > 0.4).astype(np.uint8)
    ls = np.sort(rng.integers(0, 60, 300)).astype(np.uint32)
    le = (ls + rng.integers(1, 40, 300)).astype(np.uint32)
    c = dict(rs=rs, re=re, ls=ls, le=le, scores=sc, valid=valid)
    check_all_shapes(ctx, orc, c)


def test_bits_boundary_words_across_seqnames(ctx, orc):
    """Seqname boundaries that fall inside a 32-left validity word and inside a 1024-left tile: neighbouring tiles
    share words and must not clobber each other's bits."""
    parts, idxs, off = [], [], [0]
    for s, nl in enumerate([45, 1000, 7, 2100, 0, 33]):
        c = random_case(600 + s, nl, 800, span=9000, sorted_left=True, null_frac=0.2 if s % 2 else 0.0)
        parts.append(c)
        idxs.append(ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"]) if s != 2 else None)
        off.append(off[-1] + nl)
    ls = np.concatenate([c["ls"] for c in parts])
    le = np.concatenate([c["le"] for c in parts])
    off = np.array(off, np.uint64)
    for op in ("mean", "sum", "count"):
        o_b, v_b = ctx.map_joins_batch(idxs, off, ls, le, [op])
        o, bits = ctx.map_joins_batch(idxs, off, ls, le, [op], valid_bits=True)
        got = np.unpackbits(bits.view(np.uint8), bitorder="little")[: len(ls)]
        assert (got == v_b[:, 0]).all()
        assert (o[v_b[:, 0] == 1] == o_b[v_b[:, 0] == 1]).all()
    # several columns / member reductions: packed from bytes on the device
    ops = ["mean", "min", "count", "median"]
    o_b, v_b = ctx.map_joins_batch(idxs, off, ls, le, ops)
    o, bits = ctx.map_joins_batch(idxs, off, ls, le, ops, valid_bits=True)
    got = np.unpackbits(bits.view(np.uint8), bitorder="little")[: len(ls) * 4].reshape(-1, 4)
    assert (got == v_b).all()


def test_mixed_variants_in_one_batch(ctx, orc):
    """One call over seqnames whose indexes need different kernel variants (plain, None scores, 96-bit sums,
    non-finite scores, absent): the launch is split into runs, results are those of the per-seqname calls."""
    kinds = [dict(), dict(null_frac=0.2), dict(score_kind="decimal"), dict(score_kind="nonfinite"), None, dict(score_kind="decimal", null_frac=0.3),
             dict()]
    parts, idxs, off = [], [], [0]
    for s, kw in enumerate(kinds):
        c = random_case(700 + s, 1500 + 37 * s, 2500, span=30_000, sorted_left=True, **(kw or {}))
        parts.append(c)
        idxs.append(None if kw is None else ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"]))
        off.append(off[-1] + len(c["ls"]))
    ls = np.concatenate([c["ls"] for c in parts])
    le = np.concatenate([c["le"] for c in parts])
    off = np.array(off, np.uint64)
    for ops in (["mean"], ["sum"], PIPE_OPS, ["mean", "min", "median"]):
        out, ov = ctx.map_joins_batch(idxs, off, ls, le, ops)
        for s, c in enumerate(parts):
            sl = slice(int(off[s]), int(off[s + 1]))
            if idxs[s] is None:
                want, wv = orc.map_joins(None, None, None, c["ls"], c["le"], ops)
                assert (ov[sl] == wv).all()
                continue
            want, wv, abs_sums, nvalid = oracle_refs(orc, c, ops)
            assert_map_equal(out[sl], ov[sl], want, wv, ops, abs_sums, nvalid)


def test_output_buffers_are_never_copied(ctx):
    """A caller-provided output that the library could only fill through a temporary copy is rejected."""
    c = random_case(3, 100, 100)
    ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"])
    off = np.array([0, 100], np.uint64)
    good, gv = np.empty((100, 1)), np.empty((100, 1), np.uint8)
    ctx.map_joins_batch([ix], off, c["ls"], c["le"], ["mean"], out=good, out_valid=gv)
    with pytest.raises((TypeError, ValueError)):
        ctx.map_joins_batch([ix], off, c["ls"], c["le"], ["mean"], out=np.empty((100, 1), np.float32), out_valid=gv)
    with pytest.raises((TypeError, ValueError)):
        ctx.map_joins_batch([ix], off, c["ls"], c["le"], ["mean"], out=np.empty((100, 2))[:, :1], out_valid=gv)
    with pytest.raises((TypeError, ValueError)):
        ctx.map_joins_batch([ix], off, c["ls"], c["le"], ["mean"], out=np.empty((99, 1)), out_valid=gv)


def test_unsorted_lefts_go_through_buckets(ctx, orc):
    """Lefts in any order give the reference's rows (src/granges.rs:935-984 walks the left container as it is).  Large
    unsorted calls are bucketed by coordinate on the device and their rows written back: host buffers are sampled on
    the host (AUTO), and the option forces either path -- all three must agree with the oracle and with each other."""
    import granges_b200 as gb

    c = random_case(801, 150_000, 120_000, span=3_000_000, maxlen=400, null_frac=0.05)
    nl = len(c["ls"])
    ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
    off = np.array([0, nl], np.uint64)
    for ops in (["mean"], ["sum", "count", "mean"], ["min", "median", "mean"]):
        want, wv, abs_sums, nvalid = oracle_refs(orc, c, ops)
        res = {}
        ctx.set_left_order("sorted")
        ctx.map_joins_batch([ix], off, c["ls"], c["le"], ops)  # (builds the member tables these reductions need, once)
        for order in ("auto", "sorted", "unsorted"):
            ctx.set_left_order(order)
            n0 = ctx.launch_count
            res[order] = ctx.map_joins_batch([ix], off, c["ls"], c["le"], ops) + (ctx.launch_count - n0,)
            assert_map_equal(res[order][0], res[order][1], want, wv, ops, abs_sums, nvalid)
        ctx.set_left_order("auto")
        assert res["auto"][2] == res["unsorted"][2] > res["sorted"][2]  # AUTO sampled the host buffer and bucketed
        for order in ("sorted", "unsorted"):
            assert (res[order][1] == res["auto"][1]).all() and same_bits(res[order][0][wv == 1], res["auto"][0][wv == 1])
    # sorted host lefts are left alone: AUTO launches what SORTED launches
    o = np.lexsort((c["le"], c["ls"]))
    counts = {}
    for order in ("sorted", "auto"):
        ctx.set_left_order(order)
        n0 = ctx.launch_count
        ctx.map_joins_batch([ix], off, c["ls"][o], c["le"][o], ["mean"])
        counts[order] = ctx.launch_count - n0
    ctx.set_left_order("auto")
    assert counts["auto"] == counts["sorted"]


def test_unsorted_device_lefts_are_learned(ctx, orc):
    """Device buffers in AUTO mode put nothing extra on the stream: the pipelined kernel reports the tiles it could not
    stage, and the next call on the same buffer takes the buckets.  Several seqnames, one absent, batch starting past 0."""
    import torch

    parts, idxs, off = [], [], [0]
    for s, nl in enumerate([70_000, 0, 90_000, 40_000]):
        c = random_case(810 + s, nl, 60_000, span=2_000_000, maxlen=300)
        parts.append(c)
        idxs.append(ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"]) if s != 3 else None)
        off.append(off[-1] + nl)
    ls = torch.from_numpy(np.concatenate([c["ls"] for c in parts]).astype(np.int32)).cuda()
    le = torch.from_numpy(np.concatenate([c["le"] for c in parts]).astype(np.int32)).cuda()
    off = np.array(off, np.uint64)
    launches, outs = [], []
    for rep in range(3):
        n0 = ctx.launch_count
        out, ov = ctx.map_joins_batch(idxs, off, ls, le, ["mean"])
        ctx.sync()
        launches.append(ctx.launch_count - n0)
        outs.append((out.cpu().numpy(), ov.cpu().numpy()))
    assert launches[1] > launches[0] and launches[2] == launches[1]  # learned after the first call, and stays
    for out, ov in outs:
        for s, c in enumerate(parts):
            sl = slice(int(off[s]), int(off[s + 1]))
            if idxs[s] is None:
                assert not ov[sl].any()
                continue
            want, wv, abs_sums, nvalid = oracle_refs(orc, c, ["mean"])
            assert_map_equal(out[sl], ov[sl], want, wv, ["mean"], abs_sums, nvalid)


def test_composable_reducers(ctx, orc):
    """grb_map_joins_reduce_batch: a reducer described by its parts (value, weight, fold, divide, empty) gives what the
    closure it stands for gives -- FloatOperation::run (src/data/operations.rs:53-109), the weighted mean of
    examples/weighted_mean.rs:4-31, the covered basepairs of src/commands.rs:839-843 -- and unsupported combinations fail."""
    import granges_b200 as gb

    c = random_case(901, 4000, 5000, span=90_000, maxlen=700, sorted_left=True, null_frac=0.15, score_kind="signed-decimal")
    ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
    off = np.array([0, len(c["ls"])], np.uint64)
    table = [(("score", "one", "sum", "none", "zero"), "sum"), (("score", "one", "sum", "none", "novalue"), "sum-not-empty"),
             (("score", "one", "sum", "by-weights", "novalue"), "mean"), (("one", "one", "sum", "none", "zero"), "count"),
             (("one", "overlap-width", "sum", "none", "zero"), "overlap-bp"), (("score", "overlap-width", "sum", "by-weights", "zero"), "weighted-mean"),
             (("score", "overlap-width", "sum", "none", "zero"), "weighted-sum"), (("score", "one", "min", "none", "novalue"), "min"),
             (("score", "one", "max", "none", "novalue"), "max")]
    ops = [op for _, op in table]
    out, ov = ctx.map_joins_reduce([ix], off, c["ls"], c["le"], [r for r, _ in table])
    want, wv, abs_sums, nvalid = oracle_refs(orc, c, ops)
    assert_map_equal(out, ov, want, wv, ops, abs_sums, nvalid)
    # the same numbers as the enum spelling, bit for bit
    o2, v2 = ctx.map_joins_batch([ix], off, c["ls"], c["le"], ops)
    assert (v2 == ov).all() and same_bits(o2[ov == 1], out[ov == 1])
    for bad in [("score", "overlap-width", "min", "none", "novalue"), ("one", "one", "sum", "by-weights", "zero"), ("score", "one", "min", "none", "zero")]:
        with pytest.raises(gb.GRangesError) as ei:
            ctx.map_joins_reduce([ix], off, c["ls"], c["le"], [bad])
        assert ei.value.code == -8


def test_collapse_in_sort_ranges_order(ctx, orc):
    """FloatOperation::Collapse (src/data/operations.rs:99-106) with the order the reference leaves open pinned to
    LeftGroupedJoin::sort_ranges (src/join.rs:121-128): per left, the non-missing scores of its group in (start, end, index)
    order, printed like f64::to_string and joined by ","; "" for a left without scored overlaps."""
    from tests.test_host_format import rust_display

    for seed, kw in ((31, dict(null_frac=0.3, score_kind="decimal")), (32, dict(score_kind="nonfinite")), (33, dict(score_kind="mixed", maxlen=3, span=40))):
        c = random_case(seed, 700, 900, **kw)
        ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"], c["valid"])
        rows = ctx.collapse(ix, c["ls"], c["le"])
        off, idx, _, _ = orc.left_overlaps(orc.Tree(c["rs"], c["re"]), c["ls"], c["le"])
        assert len(rows) == len(c["ls"])
        for i, row in enumerate(rows):
            members = [int(d) for d in idx[int(off[i]):int(off[i + 1])] if c["valid"] is None or c["valid"][int(d)]]
            assert row == ",".join(rust_display(float(c["scores"][d])) for d in members), i
    assert ctx.collapse(None, c["ls"][:5], c["le"][:5]) == [""] * 5
    assert ctx.collapse(ix, np.zeros(0, np.uint32), np.zeros(0, np.uint32)) == []
