"""The N > 1 path on CPU: two processes (gloo), each takes its units of the plan, filters the rights
by the unit's coordinate bounds and answers its lefts (the oracle stands in for the GPU executor --
this test checks the host logic: plan, right filter, result order), results are gathered and must
equal the single-process answer row for row."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    rng = np.random.default_rng(99)
    seqs = []
    for nl, nr, span in [(6000, 5000, 400_000), (0, 100, 1000), (2500, 9000, 300_000), (40, 0, 5000), (3000, 3000, 100_000)]:
        ls = np.sort(rng.integers(0, span, nl)).astype(np.uint32)
        le = (ls + rng.integers(1, 1000, nl)).astype(np.uint32)
        rs = rng.integers(0, span, nr).astype(np.uint32)
        re = (rs + rng.integers(1, 1000, nr)).astype(np.uint32)
        seqs.append(dict(ls=ls, le=le, rs=rs, re=re, sc=rng.random(nr)))
    return seqs


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from granges_b200 import sharding
    from oracle import oracle as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    seqs = _data()
    off = np.concatenate([[0], np.cumsum([len(s["ls"]) for s in seqs])]).astype(np.uint64)
    ls = np.concatenate([s["ls"] for s in seqs])
    le = np.concatenate([s["le"] for s in seqs])
    units = sharding.plan(off, [len(s["rs"]) for s in seqs], world, ls, le)
    mine = sharding.units_of(units, rank)
    ops = ["mean", "count", "max", "median"]
    parts = {}
    for i, u in enumerate(units):
        if u["rank"] != rank:
            continue
        s = seqs[u["seq"]]
        a, b = u["left_begin"] - int(off[u["seq"]]), u["left_end"] - int(off[u["seq"]])
        lo, hi = sharding.unit_bounds(s["ls"][a:b], s["le"][a:b])
        assert (lo, hi) == (u["right_lo"], u["right_hi"])
        m = sharding.unit_right_mask(s["rs"], s["re"], lo, hi)
        t = orc.Tree(s["rs"][m], s["re"][m])
        out, ov = orc.map_joins(t, s["sc"][m], None, s["ls"][a:b], s["le"][a:b], ops)
        parts[i] = (out, ov)
    gathered = [None] * world
    dist.all_gather_object(gathered, parts)
    nl_mine = sum(u["left_end"] - u["left_begin"] for u in mine)
    import torch

    tot = torch.tensor([nl_mine])
    dist.all_reduce(tot)
    if rank == 0:
        allparts = {}
        for g in gathered:
            allparts.update(g)
        out = sharding.concat_in_order(units, [allparts[i][0] for i in range(len(units))])
        ov = sharding.concat_in_order(units, [allparts[i][1] for i in range(len(units))])
        q.put((units, out, ov, int(tot.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharded_join_matches_single_process():
    import torch.multiprocessing as mp

    from oracle import oracle as orc

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    units, out, ov, total = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    seqs = _data()
    assert total == sum(len(s["ls"]) for s in seqs)
    assert {u["rank"] for u in units} == {0, 1}
    ops = ["mean", "count", "max", "median"]
    want = []
    for s in seqs:
        t = orc.Tree(s["rs"], s["re"]) if len(s["rs"]) else None
        want.append(orc.map_joins(t, s["sc"] if t is not None else None, None, s["ls"], s["le"], ops))
    wout = np.concatenate([w[0] for w in want])
    wov = np.concatenate([w[1] for w in want])
    assert out.shape == wout.shape and (ov == wov).all()
    m = wov == 1
    # count / max / median bit-exact; mean: the oracle sums members in tree order, which a different
    # right subset changes, hence the tolerance (the CUDA path's exact sums are split-independent)
    assert (out[:, 1:][m[:, 1:]] == wout[:, 1:][m[:, 1:]]).all()
    assert np.allclose(out[:, 0][m[:, 0]], wout[:, 0][m[:, 0]], rtol=1e-12, atol=0)
