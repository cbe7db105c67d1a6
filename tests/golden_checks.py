"""Golden-vector checks shared by the oracle (CPU) and CUDA (GPU) test files.
Each function mirrors one of the reference's own tests (file:line in
tests/golden/reference_goldens.json -> "cite")."""
from __future__ import annotations

import numpy as np


def _split(ranges):
    a = np.array(ranges, dtype=object)
    if len(a) == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0)
    s = np.array([r[0] for r in ranges], np.uint32)
    e = np.array([r[1] for r in ranges], np.uint32)
    d = np.array([r[2] if len(r) > 2 else 0.0 for r in ranges], np.float64)
    return s, e, d


def _windows_for(eng, seqlens, w):
    names = list(seqlens)
    seq, s, e = eng.windows([seqlens[n] for n in names], w["width"], w["step"], w["chop"])
    return names, seq, s, e


def check_left_overlaps_counts(eng, G):
    g = G["left_overlaps_counts"]
    names, seq, s, e = _windows_for(eng, g["seqlens"], g["windows"])
    assert len(s) == g["n_windows"]
    rs, re, sc = _split(g["right"]["chr1"])
    t = eng.build(rs, re, None, sc)
    # num_overlaps via the CSR join (left join: one group per window)
    off, idx, prs, pre = eng.left_overlaps(t, s, e)
    assert len(off) == len(s) + 1
    assert list(np.diff(off.astype(np.int64))) == g["num_overlaps"]
    # and via count_overlaps / the COUNT reduction
    assert list(eng.count(t, s, e)) == g["num_overlaps"]
    out, ov = eng.map_joins(t, s, e, ["count"])
    assert list(out[:, 0]) == g["num_overlaps"] and ov.all()


def check_map_joins_sums(eng, G):
    g = G["map_joins_sums"]
    names, seq, s, e = _windows_for(eng, g["seqlens"], g["windows"])
    assert [int(s[0]), int(e[0])] == g["first_range"]
    rs, re, sc = _split(g["right"]["chr1"])
    t = eng.build(rs, re, None, sc)
    out, ov = eng.map_joins(t, s, e, ["sum"])
    assert ov.all()
    # the reference uses assert_eq! on f64 here: exact equality
    assert list(out[:3, 0]) == g["sums_first3"]


def check_left_with_data_right_empty(eng, G):
    g = G["left_with_data_right_empty"]
    names, seq, ws, we = _windows_for(eng, g["seqlens"], g["right_windows"])
    t = eng.build(ws, we)
    ls, le, _ = _split(g["left"]["chr1"])
    off, idx, prs, pre = eng.left_overlaps(t, ls, le)
    assert len(off) - 1 == g["n_joined"]


def _filter_case(eng, G, key, anti):
    case01 = G["granges_test_case_01"]
    for c in G[key]["cases"]:
        kept = 0
        all_kept = True
        for name in case01["seqlens"]:
            ls, le, _ = _split(case01["ranges"][name])
            if name in c["right"]:
                rs, re, _ = _split(c["right"][name])
                t = eng.build(rs, re)
            else:
                t = None  # right has no such seqname
            keep = eng.filter(t, ls, le, anti)
            kept += int(keep.sum())
            all_kept &= bool(keep.all())
        assert kept == c["kept"], (key, c)
        if c.get("identity"):
            assert all_kept


def check_filter_overlaps(eng, G):
    _filter_case(eng, G, "filter_overlaps", False)


def check_antifilter_overlaps(eng, G):
    _filter_case(eng, G, "antifilter_overlaps", True)


def check_from_windows(eng, G):
    g = G["from_windows"]
    l35 = g["len35"]
    seq, s, e = eng.windows([35], l35["width"], None, True)
    assert len(s) == l35["chop_count"] and all((e - s) == l35["chop_all_width"])
    seq, s, e = eng.windows([35], l35["width"], None, False)
    assert int(e[-1] - s[-1]) == l35["nochop_last_width"]
    names = list(g["seqlens"])
    lens = [g["seqlens"][n] for n in names]
    for chop, key in ((True, "chop"), (False, "no_chop")):
        seq, s, e = eng.windows(lens, g["width"], g["step"], chop)
        got = [[names[int(q)], int(a), int(b)] for q, a, b in zip(seq, s, e)]
        assert got == g[key]
    d = G["from_windows_doctest"]
    for c in d["cases"]:
        seq, s, e = eng.windows([d["seqlens"]["chr1"]], c["width"], c["step"], c["chop"])
        got = [[int(a), int(b)] for a, b in zip(s, e)]
        assert got[: len(c["first"])] == c["first"]
        if c["step"] is None:
            assert got == c["first"]


def check_bedtools_map_example(eng, G):
    g = G["bedtools_map_example"]
    ls, le, _ = _split(g["left"]["chr1"])
    rs, re, sc = _split(g["right"]["chr1"])
    t = eng.build(rs, re, None, sc)
    out, ov = eng.map_joins(t, ls, le, ["mean"])
    for i, m in enumerate(g["mean"]):
        if m is None:
            assert ov[i, 0] == 0
        else:
            assert ov[i, 0] == 1 and out[i, 0] == m


def check_coitrees_iteration(eng, G):
    """Index order = sorted (start,end,index); observable through the canonical group order."""
    g = G["coitrees_iteration"]
    case01 = G["granges_test_case_01"]
    rs, re, sc = _split(case01["ranges"]["chr1"])
    # shuffle the input, keep the data index
    perm = np.array([2, 0, 1])
    t = eng.build(rs[perm], re[perm], perm.astype(np.uint64), sc)
    off, idx, prs, pre = eng.left_overlaps(t, np.array([0], np.uint32), np.array([30], np.uint32))
    got = [[int(a), int(b), int(i)] for a, b, i in zip(prs, pre, idx)]
    assert got == g["chr1"]
    n = sum(len(v) for v in case01["ranges"].values())
    assert n == g["len"]


def check_closed_interval(eng, G):
    """first/last of (0,10) = (0,9): a query starting at 10 must not hit, one ending at 0... cannot exist;
    a query [9,10) hits, [10,11) does not; book-ended ranges do not overlap."""
    g = G["closed_interval_conversion"]
    t = eng.build(np.array([g["range"][0]], np.uint32), np.array([g["range"][1]], np.uint32))
    ls = np.array([g["last"], g["last"] + 1], np.uint32)
    le = ls + 1
    assert list(eng.count(t, ls, le)) == [1, 0]


ALL_ENGINE_CHECKS = [
    check_left_overlaps_counts,
    check_map_joins_sums,
    check_left_with_data_right_empty,
    check_filter_overlaps,
    check_antifilter_overlaps,
    check_from_windows,
    check_bedtools_map_example,
    check_coitrees_iteration,
    check_closed_interval,
]
