"""The reference's own unit tests for the overlap-join path, restated line for line over the mirror of its container API
(granges_b200/granges.py -> C ABI -> CUDA).  Each test names the reference test it follows (paths relative to
/root/reference/src)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import __graft_entry__ as g

    g.build()
    from granges_b200 import granges

    return granges


def granges_test_case_01(G):
    """test_utilities.rs:263-268"""
    gr = G.GRanges.new_vec({"chr1": 30, "chr2": 100})
    gr.push_range("chr1", 0, 5, 1.1)
    gr.push_range("chr1", 4, 7, 8.1)
    gr.push_range("chr1", 10, 17, 10.1)
    gr.push_range("chr2", 10, 20, 3.7)
    gr.push_range("chr2", 18, 32, 1.1)
    return gr


def test_to_coitrees(G):
    """granges.rs:1704-1709"""
    gr = granges_test_case_01(G).into_coitrees()
    assert len(gr) == 5


def _keep(G, seqlens, *ranges):
    gr_keep = G.GRangesEmpty.new_vec(seqlens)
    for r in ranges:
        gr_keep.push_range(*r)
    return gr_keep.into_coitrees()


def test_granges_filter_overlaps(G):
    """granges.rs:1711-1752"""
    seqlens = {"chr1": 10}
    assert len(granges_test_case_01(G).filter_overlaps(_keep(G, seqlens, ("chr1", 0, 1)))) == 1      # test 1
    assert len(granges_test_case_01(G).filter_overlaps(_keep(G, seqlens, ("chr1", 0, 5)))) == 2      # test 2
    assert len(granges_test_case_01(G).filter_overlaps(_keep(G, seqlens, ("chr1", 8, 9)))) == 0      # test 3
    seqlens = {"chr1": 10, "chr2": 10}
    gr_filtered = granges_test_case_01(G).filter_overlaps(_keep(G, seqlens, ("chr1", 4, 7), ("chr2", 10, 12)))  # test 4
    assert len(gr_filtered) == 3
    assert list(gr_filtered.iter_ranges()) == [("chr1", 0, 5), ("chr1", 4, 7), ("chr2", 10, 20)]
    assert gr_filtered.data == [1.1, 8.1, 3.7]  # survivors keep their data, re-indexed (granges.rs:1600-1607)


def test_granges_antifilter_overlaps(G):
    """granges.rs:1754-1799"""
    seqlens = {"chr1": 10}
    assert len(granges_test_case_01(G).antifilter_overlaps(_keep(G, seqlens, ("chr1", 0, 1)))) == 4
    assert len(granges_test_case_01(G).antifilter_overlaps(_keep(G, seqlens, ("chr1", 0, 5)))) == 3
    gr = granges_test_case_01(G)
    gr_filtered = gr.antifilter_overlaps(_keep(G, seqlens, ("chr1", 8, 9)))
    assert len(gr_filtered) == 5
    assert list(gr.iter_ranges()) == list(gr_filtered.iter_ranges()) and gr.data == gr_filtered.data  # assert_eq!(gr, gr_filtered)
    seqlens = {"chr1": 10, "chr2": 10}
    assert len(granges_test_case_01(G).antifilter_overlaps(_keep(G, seqlens, ("chr1", 4, 7), ("chr2", 10, 12)))) == 2


def test_from_windows(G):
    """granges.rs:1839-1908"""
    sl = {"chr1": 35}
    chopped = G.GRangesEmpty.from_windows(sl, 10, None, True)
    assert len(chopped) == 3 and all(e - s == 10 for _, s, e in chopped.iter_ranges())
    kept = G.GRangesEmpty.from_windows(sl, 10, None, False)
    assert list(kept.iter_ranges())[-1] == ("chr1", 30, 35)
    sl = {"chr1": 21, "chr2": 13}
    gr = G.GRangesEmpty.from_windows(sl, 10, 2, True)
    assert list(gr.iter_ranges()) == [("chr1", 0, 10), ("chr1", 2, 12), ("chr1", 4, 14), ("chr1", 6, 16), ("chr1", 8, 18),
                                      ("chr1", 10, 20), ("chr2", 0, 10), ("chr2", 2, 12)]
    gr = G.GRangesEmpty.from_windows(sl, 10, 2, False)
    assert list(gr.iter_ranges())[10:] == [("chr1", 20, 21), ("chr2", 0, 10), ("chr2", 2, 12), ("chr2", 4, 13), ("chr2", 6, 13),
                                           ("chr2", 8, 13), ("chr2", 10, 13), ("chr2", 12, 13)]


def test_left_overlaps(G):
    """granges.rs:1910-1939"""
    sl = {"chr1": 50}
    windows = G.GRangesEmpty.from_windows(sl, 10, None, True)
    windows_len = len(windows)
    right_gr = G.GRanges.new_vec(sl)
    right_gr.push_range("chr1", 1, 2, 1.1)
    right_gr.push_range("chr1", 5, 7, 2.8)
    right_gr.push_range("chr1", 21, 35, 1.2)
    right_gr.push_range("chr1", 23, 24, 2.9)
    right_gr = right_gr.into_coitrees()
    joined_results = windows.left_overlaps(right_gr)
    assert len(joined_results) == windows_len  # check is left join
    assert joined_results.num_overlaps() == [2, 0, 2, 1]
    assert joined_results.overlap_widths() == [[1, 2], [], [9, 1], [5]]


def test_left_with_data_right_empty(G):
    """granges.rs:1941-1976"""
    sl = {"chr1": 50}
    left_gr = G.GRanges.new_vec(sl)
    left_gr.push_range("chr1", 1, 2, 1.1)
    left_gr.push_range("chr1", 5, 7, 2.8)
    windows = G.GRangesEmpty.from_windows(sl, 10, None, True).into_coitrees()
    assert len(left_gr.left_overlaps(windows)) == 2


def test_map_joins(G):
    """granges.rs:1978-2011 -- the closure sums the right data; also through the device-side FloatOperation::Sum."""
    sl = {"chr1": 50}
    windows = G.GRangesEmpty.from_windows(sl, 10, None, True)
    right_gr = G.GRanges.new_vec(sl)
    right_gr.push_range("chr1", 1, 2, 1.1)
    right_gr.push_range("chr1", 1, 2, 1.1)
    right_gr.push_range("chr1", 5, 7, 2.8)
    right_gr.push_range("chr1", 21, 35, 1.2)
    right_gr.push_range("chr1", 23, 24, 2.9)
    right_gr = right_gr.into_coitrees()
    joined = windows.left_overlaps(right_gr)
    assert next(iter(windows.iter_ranges())) == ("chr1", 0, 10)
    by_closure = joined.map_joins(lambda left, right_data: sum(right_data))
    assert by_closure[:3] == [5.0, 0, 4.1]
    assert joined.map_joins("sum")[:3] == [5.0, 0.0, 4.1]  # assert_eq! on f64 in the reference: exact
    assert joined.map_joins(["mean", "median", "min", "max"])[:3] == [[5.0 / 3, 1.1, 1.1, 2.8], [None] * 4, [2.05, 2.05, 1.2, 2.9]]


def test_push_range_errors(G):
    """granges.rs:731-755: a seqname outside the genome is GRangesError::MissingSequence; end <= start panics
    (ranges/mod.rs:30,109)."""
    from granges_b200 import GRangesError

    gr = G.GRangesEmpty.new_vec({"chr1": 10})
    with pytest.raises(GRangesError):
        gr.push_range("chrX", 0, 1)
    with pytest.raises(AssertionError):
        gr.push_range("chr1", 5, 5)
