"""How map results are compared with the oracle's (shared by the GPU parity tests)."""
import numpy as np

SUM_TOL = 1e-12


def assert_map_equal(out, ov, want, wv, ops, abs_sums, nvalid=None):
    """Bit-exact for min / max / median / count / overlap-bp.  sum-like columns: |error| <= 1e-12 * sum|x| over the
    group -- the bound of the reference's own left-to-right f64 sum -- and for the mean that bound divided by the
    number of members that carry a score (pass `nvalid`).  Non-finite results (+-inf, NaN) must match in kind."""
    out, want = np.asarray(out), np.asarray(want)
    assert (np.asarray(ov) == wv).all(), "validity (NoValue) mismatch"
    for c, op in enumerate(ops):
        m = wv[:, c] == 1
        if op in ("sum", "sum-not-empty", "mean", "weighted-mean", "weighted-sum"):
            o, w = out[m, c], want[m, c]
            fin = np.isfinite(w)
            assert (np.isnan(o) == np.isnan(w)).all(), (op, "NaN results differ")
            assert (o[np.isinf(w)] == w[np.isinf(w)]).all(), (op, "infinite results differ")
            tol = SUM_TOL * np.maximum(abs_sums[m], 0) + 0.0
            if op == "weighted-sum":
                tol = tol * 1e3 + SUM_TOL * np.abs(w)  # every term carries an overlap width (< maxlen of the test cases)
            if op == "mean" and nvalid is not None:
                tol = tol / np.maximum(nvalid[m], 1)
            err = np.abs(o[fin] - w[fin])
            bad = err > tol[fin]
            assert not bad.any(), (op, o[fin][bad][:5], w[fin][bad][:5])
        else:
            assert (out[m, c] == want[m, c]).all(), op


def finite_abs(scores):
    """|x| of the finite scores, 0 for the others: the scale of the rounding error of a group's sum."""
    a = np.abs(np.asarray(scores, np.float64))
    return np.where(np.isfinite(a), a, 0.0)


def same_bits(a, b):
    """Element-wise: identical f64 bit patterns, or both NaN (the payload of a NaN is not pinned)."""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    return bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())
