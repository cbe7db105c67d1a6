"""Pins the CPU oracle against every known-answer vector the reference's own
tests hold for the hot path."""
import numpy as np
import pytest

from tests import golden_checks as gc
from tests.engines import OracleEngine


@pytest.fixture(scope="module")
def eng():
    return OracleEngine()


@pytest.mark.parametrize("check", gc.ALL_ENGINE_CHECKS, ids=lambda f: f.__name__)
def test_golden(eng, goldens, check):
    check(eng, goldens)


def test_median_goldens(goldens):
    from oracle import oracle as orc

    for c in goldens["median"]["cases"]:
        assert orc.median(c["in"]) == c["out"]


def test_overlap_width_goldens(goldens):
    from oracle import oracle as orc

    for c in goldens["overlap_width"]["cases"]:
        assert orc.overlap_width(c["a"], c["b"]) == c["w"]
        assert orc.overlap_width(c["b"], c["a"]) == c["w"]


def test_float_ops_semantics():
    """data/operations.rs:53-109 corner cases."""
    from oracle import oracle as orc

    assert orc.float_op("sum", []) == 0.0
    assert orc.float_op("sum-not-empty", []) is None
    assert orc.float_op("mean", []) is None
    assert orc.float_op("min", []) is None
    assert orc.float_op("max", [float("inf"), float("nan")]) is None  # non-finite filtered
    assert orc.float_op("min", [3.0, float("-inf"), 1.0]) == 1.0
    assert orc.float_op("max", [3.0, float("inf"), 1.0]) == 3.0
    assert orc.float_op("mean", [1.0, 2.0, 6.0]) == 3.0
    assert np.isinf(orc.float_op("sum", [1.0, float("inf")]))


def test_invalid_ranges_rejected():
    """RangeIndexed::new asserts end > start (ranges/mod.rs:109); coordinates must fit i32
    (ranges/coitrees.rs:173-178)."""
    from oracle import oracle as orc

    with pytest.raises(ValueError):
        orc.Tree([5], [5])
    with pytest.raises(ValueError):
        orc.Tree([0], [2**31 + 5])
