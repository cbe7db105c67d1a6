"""BASELINE.json's full sizes, checked through size-independent properties (no CPU oracle can run
100M x 100M in test time): identities between the independent entry points, agreement between the
pipelined kernel and the one-tile-per-CTA kernel, sharded == unsharded, and an oracle spot check of
one whole seqname."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def job():
    """The bench workload: 100M x 100M synthetic hg38 ranges (bench.gen_ranges), indexes built once."""
    import torch

    import bench
    import granges_b200 as gb

    dev = torch.device("cuda", 0)
    ctx = gb.Context(0)
    n = 100_000_000
    per = bench.split_counts(n, 22)
    idxs, ls, le, rights = [], [], [], {}
    for s in range(22):
        a, b = bench.gen_ranges(s, per[s], 13, dev, False)
        ls.append(a)
        le.append(b)
        rs, re, sc = bench.gen_ranges(s, per[s], 14, dev, True)
        idxs.append(ctx.index_build(rs, re).set_scores(sc))
        if s in (13, 21):
            rights[s] = (rs.cpu().numpy().astype(np.uint32), re.cpu().numpy().astype(np.uint32), sc.cpu().numpy())
    off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
    return dict(ctx=ctx, idxs=idxs, off=off, ls=torch.cat(ls), le=torch.cat(le), n=n, rights=rights, torch=torch)


def test_full_size_identities(job):
    torch, ctx = job["torch"], job["ctx"]
    idxs, off, ls, le, n = job["idxs"], job["off"], job["ls"], job["le"], job["n"]
    mean, mv = ctx.map_joins_batch(idxs, off, ls, le, ["mean"])
    ssum, sv = ctx.map_joins_batch(idxs, off, ls, le, ["sum"])
    cnt_f, _ = ctx.map_joins_batch(idxs, off, ls, le, ["count"])
    cnt = ctx.count_overlaps_batch(idxs, off, ls, le)
    keep, nk = ctx.filter_overlaps_batch(idxs, off, ls, le)
    ctx.sync()
    c = cnt.to(torch.int64)
    assert int(c.sum().item()) == 4_327_288_487  # the workload's overlap pairs (also what bench.py reports)
    assert (cnt_f[:, 0].to(torch.int64) == c).all()
    assert (keep.to(torch.bool) == (c > 0)).all() and nk == int((c > 0).sum().item())
    assert (mv[:, 0].to(torch.bool) == (c > 0)).all() and sv.all()
    nz = c > 0
    # mean is sum / count with IEEE division of the same exactly-rounded sum: bit-identical
    assert (mean[:, 0][nz] == ssum[:, 0][nz] / c[nz].to(torch.float64)).all()
    assert (ssum[:, 0][~nz] == 0).all()
    # scores are in [0, 1): 0 <= mean < 1
    assert float(mean[:, 0][nz].min()) >= 0.0 and float(mean[:, 0][nz].max()) < 1.0


def test_full_size_kernels_agree(job, monkeypatch):
    """k_map_pipe (persistent TMA pipeline) and k_tile (one tile per CTA) must agree bit for bit."""
    import granges_b200 as gb

    ctx, idxs, off, ls, le = job["ctx"], job["idxs"], job["off"], job["ls"], job["le"]
    a, av = ctx.map_joins_batch(idxs, off, ls, le, ["mean"])
    monkeypatch.setenv("GRB_PIPE_MODE", "0")
    ctx2 = gb.Context(0)
    # indexes belong to their context: rebuild two seqnames under the other context and compare those rows
    import bench

    torch = job["torch"]
    for s in (0, 20):
        rs, re, sc = bench.gen_ranges(s, int(off[s + 1] - off[s]), 14, torch.device("cuda", 0), True)
        ix = ctx2.index_build(rs, re).set_scores(sc)
        sl = slice(int(off[s]), int(off[s + 1]))
        b, bv = ctx2.map_joins(ix, ls[sl], le[sl], ["mean"])
        ctx2.sync()
        ctx.sync()
        assert (bv == av[sl]).all() and (b[bv.bool()] == a[sl][bv.bool()]).all()


def test_full_size_oracle_spot_check(job):
    """One whole seqname (4.5M x 4.5M) of the 100M x 100M result against the oracle: counts exact, mean 1e-12."""
    from oracle import oracle as orc

    ctx, idxs, off, ls, le = job["ctx"], job["idxs"], job["off"], job["ls"], job["le"]
    out, ov = ctx.map_joins_batch(idxs, off, ls, le, ["mean", "count"])
    ctx.sync()
    for s in (13, 21):
        rs, re, sc = job["rights"][s]
        sl = slice(int(off[s]), int(off[s + 1]))
        hl = ls[sl].cpu().numpy().astype(np.uint32)
        he = le[sl].cpu().numpy().astype(np.uint32)
        want, wv = orc.map_joins(orc.Tree(rs, re), sc, None, hl, he, ["mean", "count"], nthreads=8)
        got, gv = out[sl].cpu().numpy(), ov[sl].cpu().numpy()
        assert (gv == wv).all()
        assert (got[:, 1] == want[:, 1]).all()
        m = wv[:, 0] == 1
        assert np.allclose(got[m, 0], want[m, 0], rtol=1e-12, atol=0)


def test_full_size_member_reductions_vs_oracle(job):
    """min / max / median (config C3's member sweep: k_sweep3, and k_minmax without a median beside them) on one WHOLE
    seqname of the 100M x 100M job against the oracle: bit-exact."""
    from oracle import oracle as orc

    ctx, idxs, off, ls, le = job["ctx"], job["idxs"], job["off"], job["ls"], job["le"]
    s = 21
    rs, re, sc = job["rights"][s]
    sl = slice(int(off[s]), int(off[s + 1]))
    hl = ls[sl].cpu().numpy().astype(np.uint32)
    he = le[sl].cpu().numpy().astype(np.uint32)
    tree = orc.Tree(rs, re)
    for ops in (["min", "max", "mean", "sum", "median"], ["min", "max"], ["median"]):
        out, ov = ctx.map_joins_batch(idxs, off, ls, le, ops)
        ctx.sync()
        want, wv = orc.map_joins(tree, sc, None, hl, he, ops, nthreads=8)
        got, gv = out[sl].cpu().numpy(), ov[sl].cpu().numpy()
        assert (gv == wv).all()
        for c, op in enumerate(ops):
            m = wv[:, c] == 1
            if op in ("mean", "sum"):
                assert np.allclose(got[m, c], want[m, c], rtol=1e-12, atol=0)
            else:
                assert (got[m, c] == want[m, c]).all(), op
        del out, ov


def test_full_size_clustered_heavy_groups_vs_oracle():
    """Config C4's right set on the seqname with the heaviest groups (chr21: 22.7M clustered BED5 rights, 80 % of them in
    two 1 Mb bins) under 1 kb / 500 bp windows: groups of more than 20 000 members go through the radix-select path
    (k_median_heavy3) and whole-warp groups; mean / min / max / median / count against the oracle."""
    import torch

    import bench
    import granges_b200 as gb
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    ctx = gb.Context(0)
    s = 20
    nr = bench.split_counts(500_000_000, 22)[s]
    rs, re, sc = bench.clustered_rights(s, nr, 15, dev)
    ix = ctx.index_build(rs, re).set_scores(sc)
    L = bench.HG38[s]
    ws = torch.arange(0, L, 500, dtype=torch.int64, device=dev)
    we = torch.clamp(ws + 1000, max=L)
    ws, we = ws.to(torch.int32), we.to(torch.int32)
    ops = ["mean", "min", "max", "median", "count"]
    out, ov = ctx.map_joins(ix, ws, we, ops)
    ctx.sync()
    want, wv = orc.map_joins(orc.Tree(rs.cpu().numpy().astype(np.uint32), re.cpu().numpy().astype(np.uint32)), sc.cpu().numpy(), None,
                             ws.cpu().numpy().astype(np.uint32), we.cpu().numpy().astype(np.uint32), ops, nthreads=8)
    got, gv = out.cpu().numpy(), ov.cpu().numpy()
    assert want[:, 4].max() > 20_000  # the heavy-group paths are exercised
    assert (gv == wv).all()
    for c, op in enumerate(ops):
        m = wv[:, c] == 1
        if op == "mean":
            assert np.allclose(got[m, c], want[m, c], rtol=1e-12, atol=0)
        else:
            assert (got[m, c] == want[m, c]).all(), op


def test_full_size_pairs_vs_oracle():
    """Config C5 (left_overlaps with every pair materialised) on one seqname at its full size -- chr1: 2.27M GTF-like
    lefts x 9.09M reads-like rights -- for the first 150k lefts: offsets, right indices, right starts / ends bit for bit
    (groups in sort_ranges order, src/join.rs:121-128); and the pair count of the whole seqname against count_overlaps."""
    import torch

    import bench
    import granges_b200 as gb
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    ctx = gb.Context(0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    nl, nr = bench.split_counts(50_000_000, 22)[0], bench.split_counts(200_000_000, 22)[0]
    ls, le = bench.gtf_like_lefts(0, nl, g, dev)
    rs, re = bench.reads_like_rights(0, nr, g, dev)
    ix = ctx.index_build(rs, re)
    nchk = 150_000
    hl, he = ls[:nchk].cpu().numpy().astype(np.uint32), le[:nchk].cpu().numpy().astype(np.uint32)
    tree = orc.Tree(rs.cpu().numpy().astype(np.uint32), re.cpu().numpy().astype(np.uint32))
    woff, widx, wrs, wre = orc.left_overlaps(tree, hl, he)
    goff, gidx, grs, gre = ctx.left_overlaps(ix, ls[:nchk].contiguous(), le[:nchk].contiguous())
    assert (goff == woff).all() and (gidx == widx).all() and (grs == wrs).all() and (gre == wre).all()
    assert int(woff[-1]) > 10_000_000
    # the whole seqname: the CSR offsets of the materialised join end at the sum of count_overlaps
    import ctypes as C

    from granges_b200 import _capi

    off = torch.empty(nl + 1, dtype=torch.int64, device=dev)
    ph = C.c_void_p()
    rc = _capi.lib().grb_left_overlaps(ctx.h, ix.h, ls.data_ptr(), le.data_ptr(), nl, off.data_ptr(), C.byref(ph))
    assert rc == 0
    npairs = _capi.lib().grb_pairs_len(ph)
    _capi.lib().grb_pairs_destroy(ph)
    cnt = ctx.count_overlaps(ix, ls, le)
    ctx.sync()
    assert npairs == int(cnt.to(torch.int64).sum().item()) == int(off[-1].item())
    assert (torch.diff(off) == cnt.to(torch.int64)).all()


def test_full_size_whole_call_on_every_gpu_count(job):
    """`granges map` in one call from host arrays at the full 100M x 100M size: grb_map_joins_whole_f64, and
    grb_map_joins_whole_f64_mgpu on every GPU count the box offers, return the bits of the device-resident batch call."""
    import bench
    import granges_b200 as gb

    torch = job["torch"]
    ctx, idxs, off, ls, le, n = job["ctx"], job["idxs"], job["off"], job["ls"], job["le"], job["n"]
    want, wv = ctx.map_joins_batch(idxs, off, ls, le, ["mean"])
    ctx.sync()
    want, wv = want.cpu().numpy(), wv.cpu().numpy()
    per = bench.split_counts(n, 22)
    rights = []
    for s in range(22):
        rs, re, sc = bench.gen_ranges(s, per[s], 14, torch.device("cuda", 0), True)
        rights.append((rs.cpu().numpy().view(np.uint32), re.cpu().numpy().view(np.uint32), sc.cpu().numpy()))
    hl, he = ls.cpu().numpy().view(np.uint32), le.cpu().numpy().view(np.uint32)
    ctx2 = gb.Context(0)
    got, gv = ctx2.map_whole(rights, off, hl, he, ["mean"])
    assert (gv == wv).all() and (got[wv == 1] == want[wv == 1]).all()
    ndev = torch.cuda.device_count()
    for nd in sorted({1, min(2, ndev), ndev}):
        got, gv = gb.map_whole_mgpu(list(range(nd)), rights, off, hl, he, ["mean"])
        assert (gv == wv).all() and (got[wv == 1] == want[wv == 1]).all(), nd


def test_full_size_sharded_equals_unsharded(job):
    """A 3-way coordinate split of one seqname (what bench.py --gpus N does) reproduces the unsplit bits."""
    import bench
    from granges_b200 import sharding

    torch = job["torch"]
    ctx, off, ls, le = job["ctx"], job["off"], job["ls"], job["le"]
    s = 20
    sl = slice(int(off[s]), int(off[s + 1]))
    whole, wv = ctx.map_joins(job["idxs"][s], ls[sl], le[sl], ["mean"])
    rs, re, sc = bench.gen_ranges(s, int(off[s + 1] - off[s]), 14, torch.device("cuda", 0), True)
    n = sl.stop - sl.start
    cuts = [0, n // 3, 2 * n // 3, n]
    for a, b in zip(cuts, cuts[1:]):
        l_s, l_e = ls[sl][a:b].contiguous(), le[sl][a:b].contiguous()
        lo, hi = sharding.unit_bounds(l_s, l_e)
        m = sharding.unit_right_mask(rs, re, lo, hi)
        ix = ctx.index_build(rs[m].contiguous(), re[m].contiguous()).set_scores(sc[m].contiguous())
        o, v = ctx.map_joins(ix, l_s, l_e, ["mean"])
        ctx.sync()
        assert (v == wv[a:b]).all() and (o[v.bool()] == whole[a:b][v.bool()]).all()
