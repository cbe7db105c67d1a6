"""CPU model of k_sweep3's wavelet runs (tools/wavelet_median_model.py: slice, dense local ranks, the C ++ end-order range ++ R
sequence, levelwise trees with arithmetic node boundaries, one rank query per level and tree) against a brute force over the
members (src/data/operations.rs:17-32: the median is an order statistic of the group's scores)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import wavelet_median_model as wm  # noqa: E402


def test_model_matches_brute_force():
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(60):
        n = int(rng.integers(1, 300))
        nl = int(rng.integers(1, 40))
        checked += wm.run_case(rng, n, nl, int(rng.integers(2, 250)), int(rng.integers(50, 2500)))
    assert checked > 1000


def test_model_dense_and_sparse_extremes():
    rng = np.random.default_rng(12)
    # heavy coverage (every left overlaps most rights) and almost none
    assert wm.run_case(rng, 250, 30, 2000, 300) > 0
    wm.run_case(rng, 40, 30, 3, 100000)
