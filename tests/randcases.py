"""Seeded random per-seqname cases shared by oracle-vs-brute (CPU) and CUDA-vs-oracle (GPU) tests."""
from __future__ import annotations

import numpy as np


def random_case(seed, nl, nr, span=2000, maxlen=60, null_frac=0.0, score_kind="uniform", long_frac=0.0,
                shuffle_right=True, sorted_left=False):
    rng = np.random.default_rng(seed)

    def ranges(n):
        ln = rng.integers(1, maxlen + 1, n)
        if long_frac > 0:
            big = rng.random(n) < long_frac
            ln = np.where(big, rng.integers(maxlen, span, n), ln)
        ln = np.minimum(ln, span)
        st = (rng.random(n) * (span - ln + 1)).astype(np.int64)
        return st.astype(np.uint32), (st + ln).astype(np.uint32)

    ls, le = ranges(nl)
    rs, re = ranges(nr)
    if sorted_left:
        o = np.lexsort((le, ls))
        ls, le = ls[o], le[o]
    if not shuffle_right:
        o = np.lexsort((re, rs))
        rs, re = rs[o], re[o]
    if score_kind == "uniform":
        sc = rng.random(nr)
    elif score_kind == "int":
        sc = rng.integers(-50, 1000, nr).astype(np.float64)
    elif score_kind == "mixed":
        sc = rng.standard_normal(nr) * 10.0 ** rng.integers(-3, 4, nr)
    elif score_kind == "wide":  # dynamic range too wide for the exact fixed-point path
        sc = rng.standard_normal(nr) * 10.0 ** rng.integers(-30, 30, nr)
    elif score_kind == "decimal":  # what BED files hold: decimal fractions over a few decades (about 70 bits of fixed point)
        sc = np.round(0.001 + rng.random(nr) * 456.788, 3)
    elif score_kind == "wide96":  # 53-bit fractions over 31 binary decades: 84 bits, groups beyond 2^11 members leave the fast path
        sc = rng.random(nr) * 2.0 ** (-rng.integers(0, 31, nr))
    elif score_kind == "signed-decimal":  # mixed signs: cancellation
        sc = np.round((rng.random(nr) - 0.5) * 2000.0, 2)
    elif score_kind == "nan":
        sc = np.round(rng.random(nr) * 10, 2)
        if nr:
            sc[rng.integers(0, nr, max(1, nr // 50))] = np.nan
            sc[rng.integers(0, nr, max(1, nr // 50))] = np.inf
    elif score_kind == "nonfinite":
        sc = rng.random(nr)
        if nr:
            sc[rng.integers(0, nr, max(1, nr // 10))] = np.inf
            sc[rng.integers(0, nr, max(1, nr // 20))] = -np.inf
    else:
        raise ValueError(score_kind)
    valid = None
    if null_frac > 0:
        valid = (rng.random(nr) >= null_frac).astype(np.uint8)
    return dict(ls=ls, le=le, rs=rs, re=re, scores=sc, valid=valid)


ALL_OPS = ["sum", "sum-not-empty", "min", "max", "mean", "median", "count", "overlap-bp", "weighted-mean", "weighted-sum"]
