"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/granges_b200.h declares (no compute calls: there is no GPU here), the Python signature
table covers the header, and the host-side shard planner behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "granges_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from granges_b200 import _capi

    return _capi


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/granges_b200.h but not exported"


def test_python_signature_table_matches_header(built):
    assert sorted(built.SIGNATURES) == _header_functions()


def test_abi_version(built):
    assert built.lib().grb_abi_version() == 1


def test_no_cpu_fallback_without_gpu(built):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import granges_b200 as gb

    with pytest.raises(gb.GRangesError) as ei:
        gb.Context(0)
    assert ei.value.code == -4


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "granges_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|liboracle|oracle/", txt), f"{f} reaches into oracle/"


def test_shard_plan_covers_everything(built):
    import granges_b200 as gb

    rng = np.random.default_rng(3)
    nls = [1000, 0, 5000, 20, 3000]
    off = np.concatenate([[0], np.cumsum(nls)]).astype(np.uint64)
    ls = np.concatenate([np.sort(rng.integers(0, 100000, n)) for n in nls]).astype(np.uint32)
    le = ls + rng.integers(1, 500, len(ls)).astype(np.uint32)
    nr = [900, 10, 7000, 0, 100]
    for nranks in (1, 2, 3, 8):
        plan = gb.shard_plan(off, nr, ls, le, nranks)
        assert [p["rank"] for p in plan] == sorted(p["rank"] for p in plan)
        assert max(p["rank"] for p in plan) <= nranks - 1
        for s, n in enumerate(nls):
            pieces = [p for p in plan if p["seq"] == s]
            if n == 0:
                assert not pieces
                continue
            assert pieces[0]["left_begin"] == off[s] and pieces[-1]["left_end"] == off[s + 1]
            for a, b in zip(pieces, pieces[1:]):
                assert a["left_end"] == b["left_begin"]
            for p in pieces:
                sl = slice(p["left_begin"], p["left_end"])
                assert p["right_lo"] == ls[sl].min() and p["right_hi"] == le[sl].max()
        cost = np.zeros(nranks)
        for p in plan:
            s = p["seq"]
            cost[p["rank"]] += (p["left_end"] - p["left_begin"]) * (nls[s] + nr[s]) / nls[s]
        if nranks > 1:
            assert cost.max() <= 1.25 * cost.sum() / nranks


def test_pair_aware_shard_cost_balances_member_sweeps(built):
    """bench.py --ops min,...,median adds PAIR_WEIGHT x expected pairs to a seqname's cost (sharding.PAIR_WEIGHT): with
    equal counts per seqname the short chromosomes hold several times the pairs of the long ones, and the plan must
    balance lefts + weighted pairs, not lefts + rights."""
    import bench
    from granges_b200 import sharding

    n = 100_000_000
    per = bench.split_counts(n, 22)
    off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
    pairs = [per[s] * per[s] * 1000.0 / bench.HG38[s] for s in range(22)]
    cost = [int(per[s] + sharding.PAIR_WEIGHT * pairs[s]) for s in range(22)]
    for world in (2, 4, 8):
        plan = sharding.plan(off, cost, world)
        work = np.zeros(world)
        for u in plan:
            s = u["seq"]
            frac = (u["left_end"] - u["left_begin"]) / per[s]
            work[u["rank"]] += frac * (per[s] + per[s] + sharding.PAIR_WEIGHT * pairs[s])
        assert work.max() <= 1.02 * work.mean(), (world, work)
        # the lefts + rights plan of the same job is off by tens of percent on that measure
        plain = sharding.plan(off, per, world)
        w2 = np.zeros(world)
        for u in plain:
            s = u["seq"]
            frac = (u["left_end"] - u["left_begin"]) / per[s]
            w2[u["rank"]] += frac * (per[s] + per[s] + sharding.PAIR_WEIGHT * pairs[s])
        assert w2.max() > 1.2 * w2.mean()
