"""CUDA merge (grb_merge_ranges / grb_merged_map_f64) against the oracle's restatement of the merging
iterators: the reference's goldens, sorted / unsorted / dense / sparse record streams, every distance sign,
both device paths (prefix-max scan and the one-warp-per-run state machine), and a size-independent
property at 20M records."""
import numpy as np
import pytest

from tests.mergecases import merge_case, run_offsets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from tests.engines import CudaEngine

    return CudaEngine()


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


def check(eng, orc, c, distance, ops=("sum", "mean", "min", "max", "median", "sum-not-empty")):
    ro = run_offsets(c["seq"])
    want = orc.merge(c["seq"], c["starts"], c["ends"], distance)
    rom, ms, me, first, out, ov = eng.ctx.merge(c["starts"], c["ends"], ro, distance, c["scores"], c["valid"], list(ops))
    assert len(ms) == len(want[1])
    assert (ms == want[1]).all() and (me == want[2]).all() and (first[:-1] == want[3]).all() and first[-1] == len(c["starts"])
    # merged ranges per run: the seqname of merged range g is the seqname of its first record
    assert rom[-1] == len(ms) and (np.diff(rom.astype(np.int64)) >= 0).all()
    run_of_group = np.searchsorted(ro, first[:-1], side="right") - 1
    assert (np.searchsorted(rom, np.arange(len(ms)), side="right") - 1 == run_of_group).all()
    abs_sum = orc.merge(c["seq"], c["starts"], c["ends"], distance, np.abs(c["scores"]), c["valid"], "sum")[4]
    for col, op in enumerate(ops):
        w = orc.merge(c["seq"], c["starts"], c["ends"], distance, c["scores"], c["valid"], op)
        assert (ov[:, col] == w[5]).all(), op
        m = w[5] == 1
        if op in ("sum", "mean", "sum-not-empty"):
            assert (np.abs(out[m, col] - w[4][m]) <= 1e-12 * abs_sum[m]).all(), op
        else:
            assert (out[m, col] == w[4][m]).all(), op


def test_merge_goldens(eng, goldens):
    G = goldens["merging_iterators"]
    recs = G["records"]
    names = sorted({r[0] for r in recs})
    seq = np.array([names.index(r[0]) for r in recs], np.uint32)
    s, e = np.array([r[1] for r in recs], np.uint32), np.array([r[2] for r in recs], np.uint32)
    sc = np.array([r[3] for r in recs])
    ro = run_offsets(seq)
    for case in G["cases"]:
        res = eng.ctx.merge(s, e, ro, case["distance"], sc, None, [case["op"]] if "op" in case else None)
        rom, ms, me, first = res[:4]
        gseq = seq[first[:-1]]
        got = [[names[a], int(b), int(c)] for a, b, c in zip(gseq, ms, me)]
        assert got == [m[:3] for m in case["merged"]], case["cite"]
        if "op" in case:
            vals = res[4][:, 0]
            for v, m in zip(vals, case["merged"]):
                assert abs(v - m[3]) <= 1e-12 * abs(m[3])  # exact sum, rounded once, vs the reference's left-to-right f64 sum


@pytest.mark.parametrize("distance", [0, 1, 25, 1000, -1, -7, -60])
@pytest.mark.parametrize("kind", ["sorted", "unsorted", "dense", "sparse"])
def test_merge_vs_oracle(eng, orc, kind, distance):
    check(eng, orc, merge_case(500 + distance, kind, 20000), distance)


def test_merge_paths_agree(eng, orc, monkeypatch):
    """Sorted input, distance >= 0: the scan path and the state machine give the same groups."""
    c = merge_case(77, "sorted", 50000, nseq=7)
    ro = run_offsets(c["seq"])
    a = eng.ctx.merge(c["starts"], c["ends"], ro, 3)
    monkeypatch.setenv("GRB_MERGE_MODE", "1")
    b = eng.ctx.merge(c["starts"], c["ends"], ro, 3)
    for x, y in zip(a, b):
        assert (x == y).all()


def test_merge_edge_cases(eng, orc):
    # empty input, a single record, empty runs between runs, everything merging into one range
    rom, ms, me, first = eng.ctx.merge(np.zeros(0, np.uint32), np.zeros(0, np.uint32), [0], 0)
    assert len(ms) == 0 and first.tolist() == [0]
    rom, ms, me, first = eng.ctx.merge([5], [9], None, 0)
    assert ms.tolist() == [5] and me.tolist() == [9] and first.tolist() == [0, 1]
    s = np.array([1, 2, 3, 10, 11], np.uint32)
    e = s + 5
    rom, ms, me, first = eng.ctx.merge(s, e, [0, 3, 3, 5], 0)
    assert ms.tolist() == [1, 10] and me.tolist() == [8, 16] and rom.tolist() == [0, 1, 1, 2]
    n = 5000
    s = np.arange(n, dtype=np.uint32) * 3
    res = eng.ctx.merge(s, s + 3, None, 0, np.ones(n), None, ["sum", "median"])
    assert res[1].tolist() == [0] and res[2].tolist() == [3 * n] and res[4].tolist() == [[float(n), 1.0]]
    # book-ended ranges merge at distance 0 but not at -1
    assert len(eng.ctx.merge(s, s + 3, None, -1)[1]) == n
    with pytest.raises(Exception):
        eng.ctx.merge([5], [5], None, 0)


def test_merge_large_property(eng):
    """20M sorted records: merged ranges are disjoint, further apart than the distance, cover every record,
    and the per-group counts add up."""
    rng = np.random.default_rng(3)
    n, d = 20_000_000, 2
    st = np.sort(rng.integers(0, 2_000_000_000, n)).astype(np.uint32)
    en = st + rng.integers(1, 60, n).astype(np.uint32)
    rom, ms, me, first, out, ov = eng.ctx.merge(st, en, None, d, np.ones(n), None, ["sum"])
    assert (ms[1:].astype(np.int64) - me[:-1].astype(np.int64) > d).all()
    assert (ms == st[first[:-1]]).all()
    assert (np.maximum.reduceat(en, first[:-1].astype(np.int64)) == me).all()
    assert (out[:, 0] == np.diff(first.astype(np.int64))).all() and int(out[:, 0].sum()) == n


def test_merge_at_the_coordinate_limits(eng, orc):
    top = 2**31 - 1
    s = np.array([0, 10, top - 50, top - 40, top - 5], np.uint32)
    e = np.array([5, top - 45, top - 45, top - 10, top], np.uint32)
    seq = np.zeros(5, np.uint32)
    for d in (0, 5, 2**31, -1, -(2**31)):
        want = orc.merge(seq, s, e, d)
        rom, ms, me, first = eng.ctx.merge(s, e, None, d)
        assert (ms == want[1]).all() and (me == want[2]).all() and (first[:-1] == want[3]).all(), d
