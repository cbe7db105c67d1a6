#!/usr/bin/env python
"""bench.py -- `granges map --func mean` hot path (left_overlaps + map_joins(mean)) on synthetic hg38.

Workload (BASELINE.json north_star / metric): 100M left x 100M right ranges over the 22 hg38
autosomes, generated like the reference's random_granges (src/test_utilities.rs:60-164: lengths
U{1..999}, starts uniform, scores U[0,1), equal share per seqname), both sides sorted.

  value      left ranges/s of one map pass (left_overlaps + map_joins(mean), index already built,
             everything resident in HBM), CUDA events on the launching stream, max over ranks
  e2e        the same metric through ONE call of the C ABI from HOST buffers: grb_map_joins_whole_f64 (N = 1) or
             grb_map_joins_whole_f64_mgpu (N > 1, one process driving the N GPUs): H2D of rights, scores and
             lefts, index builds, map pass, D2H of the results into one output array, all inside the timed
             region; page-locked buffers for the headline figure, pageable ones beside it
  variants   the headline job on the other input classes (None scores, decimal scores, a non-finite score,
             shuffled lefts) with an oracle spot check each;  configs: BASELINE configs C2-C5 likewise
  roofline   the dominant kernel (k_map_pipe<MEAN>: the only kernel of a map_mean step) against the measured HBM peak
  cpu_baseline / --impl reference
             the CPU oracle (coitrees-restated, oracle/granges_oracle.c) on a bounded sample
             (one whole seqname) -- the reference is Rust and cannot be built in this image
N > 1: strong scaling -- the same 100M x 100M job split into independent units by seqname /
coordinate (grb_shard_plan), no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
        135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983,
        50818468]  # chr1..chr22, tests_data/hg38_seqlens.tsv

M63 = (1 << 63) - 1


def _lsr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def _mix(x):
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic): identical on CPU and CUDA."""
    z = x + (-7046029254386353131)
    z = (z ^ _lsr(z, 30)) * (-4658895280553007687)
    z = (z ^ _lsr(z, 27)) * (-7723592293110705685)
    return z ^ _lsr(z, 31)


def gen_ranges(seq, n, seed, device, with_scores):
    """n sorted ranges on seqname `seq`: (starts int32, ends int32[, scores f64 by sorted position])."""
    L = HG38[seq]
    i = torch.arange(n, dtype=torch.int64, device=device)
    base = i * 4 + (seed * 1000003 + seq * 7919) * 1_000_000_007
    ln = 1 + (_mix(base) & M63) % 999
    st = (_mix(base + 1) & M63) % (L - ln + 1)
    key, _ = torch.sort(st * 4294967296 + (st + ln))
    starts = (key >> 32).to(torch.int32)
    ends = (key & 0xFFFFFFFF).to(torch.int32)
    if not with_scores:
        return starts, ends
    sc = _lsr(_mix(base + 2), 11).to(torch.float64) * (2.0 ** -53)
    return starts, ends, sc


def split_counts(total, nseq):
    q, r = divmod(total, nseq)
    return [q + (1 if s < r else 0) for s in range(nseq)]


def clustered_rights(seq, n, seed, dev):
    """Config C4's right set: 80 % of the ranges in 5 % of the 1 Mb bins (log-normal bin weights, sigma 1.5),
    the rest uniform; len U{1..999}, scores U[0,1); sorted.  torch's generator: device-specific streams, so the CPU
    checks copy the arrays back instead of regenerating them."""
    L = HG38[seq]
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1000 + seq)
    nb = (L + 999_999) // 1_000_000
    hot = torch.randperm(nb, generator=g, device=dev)[: max(1, nb // 20)]
    w = torch.exp(1.5 * torch.randn(len(hot), generator=g, device=dev))
    nh = int(n * 0.8)
    which = hot[torch.multinomial(w, nh, replacement=True, generator=g)]
    ln = torch.randint(1, 1000, (n,), generator=g, device=dev)
    st_hot = which * 1_000_000 + torch.randint(0, 1_000_000, (nh,), generator=g, device=dev)
    st_uni = (torch.rand(n - nh, generator=g, device=dev, dtype=torch.float64) * L).to(torch.int64)
    st = torch.cat([st_hot, st_uni])
    st = torch.minimum(st, torch.tensor(L, device=dev) - ln)
    st = torch.clamp(st, min=0)
    key, _ = torch.sort(st * 4294967296 + (st + ln))
    sc = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    return (key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32), sc


def gtf_like_lefts(seq, n, g, dev):
    """Config C5's lefts: len ~ log-normal(median 2 kb, sigma 1.0) clipped to [50, 2e6], starts uniform; sorted."""
    L = HG38[seq]
    ln = torch.exp(np.log(2000.0) + torch.randn(n, generator=g, device=dev, dtype=torch.float64)).clamp(50, 2e6).to(torch.int64)
    st = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (L - ln).to(torch.float64)).to(torch.int64)
    key, _ = torch.sort(st * 4294967296 + (st + ln))
    return (key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32)


def reads_like_rights(seq, n, g, dev):
    """Config C5's rights: len U{75..151}, starts uniform; sorted."""
    L = HG38[seq]
    rl = torch.randint(75, 152, (n,), generator=g, device=dev)
    rst = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (L - 151)).to(torch.int64)
    key, _ = torch.sort(rst * 4294967296 + (rst + rl))
    return (key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32)


def variant_scores(kind, seq, n, dev):
    """Scores (and validity) of the input classes the reference takes at one speed: `nulls` = 1 % BED "." scores,
    `wide` = decimal fractions over five decades (about 70 bits of fixed point: the 96-bit prefix planes),
    `nonfinite` = one +inf per seqname (prefix counters + the IEEE rule).  Same on CPU and CUDA."""
    i = torch.arange(n, dtype=torch.int64, device=dev)
    base = i * 4 + (14 * 1000003 + seq * 7919) * 1_000_000_007
    sc = _lsr(_mix(base + 2), 11).to(torch.float64) * (2.0 ** -53)
    valid = None
    if kind == "nulls":
        valid = ((_mix(base + 3) & M63) % 100 != 0).to(torch.uint8)
    elif kind == "wide":
        sc = torch.round((0.001 + sc * 456.788) * 1000.0) / 1000.0
    elif kind == "nonfinite":
        sc = sc.clone()
        sc[n // 2] = float("inf")
    return sc, valid


def event_ms(fn, steps, warmup):
    """Mean CUDA-event time of fn() on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def map_matches_oracle(got, gv, want, wv, ops):
    """The parity rule of the tests: validity equal; min / max / median / count bit-exact; sum / mean within 1e-12
    relative; non-finite results equal in kind."""
    if not (gv == wv).all():
        return False
    for c, op in enumerate(ops):
        m = wv[:, c] == 1
        g, w = got[m, c], want[m, c]
        if op in ("sum", "mean", "sum-not-empty"):
            fin = np.isfinite(w)
            if not ((np.isnan(g) == np.isnan(w)).all() and (g[np.isinf(w)] == w[np.isinf(w)]).all()):
                return False
            if fin.any() and float(np.max(np.abs(g[fin] - w[fin]) / np.maximum(np.abs(w[fin]), 1e-300))) > 1e-12:
                return False
        elif not (g == w).all():
            return False
    return True


CHECK_SEQ, CHECK_LEFTS = 20, 200_000  # oracle spot checks: the first 200k lefts of chr21 against all of chr21's rights


def run_variants(ctx, idxs, offs, ls, le, nl_per, nr_per, dev, base_ms, steps, warmup):
    """The headline job (map mean, 100M x 100M) on the input classes that used to leave the pipelined kernel, plus
    shuffled lefts: ms per pass, ratio to the sorted / all-Some pass, oracle spot check.  Leaves the indexes with their
    original scores."""
    from oracle import oracle as orc

    s = CHECK_SEQ
    rs_h, re_h, _ = (t.numpy() for t in gen_ranges(s, nr_per[s], 14, "cpu", True))
    tree = orc.Tree(rs_h.astype(np.uint32), re_h.astype(np.uint32))
    a = int(offs[s])
    nchk = min(CHECK_LEFTS, nl_per[s])
    k = 1
    out = torch.empty((len(ls), k), dtype=torch.float64, device=dev)
    ov = torch.empty((len(ls), k), dtype=torch.uint8, device=dev)
    res = {}

    def check(ls_d, le_d, rows, sc_h, va_h):
        hl = ls_d[rows].cpu().numpy().astype(np.uint32)
        he = le_d[rows].cpu().numpy().astype(np.uint32)
        want, wv = orc.map_joins(tree, sc_h, va_h, hl, he, ["mean"], nthreads=orc.max_threads())
        return map_matches_oracle(out[rows].cpu().numpy(), ov[rows].cpu().numpy(), want, wv, ["mean"])

    for kind in ("nulls", "wide", "nonfinite"):
        for q, ix in enumerate(idxs):
            sc, va = variant_scores(kind, q, nr_per[q], dev)
            ix.set_scores(sc, va)
        call = ctx.prepare_map(idxs, offs, ls, le, ["mean"], out, ov)
        ms = event_ms(call, steps, warmup)
        sc_h, va_h = variant_scores(kind, s, nr_per[s], "cpu")
        ok = check(ls, le, slice(a, a + nchk), sc_h.numpy(), None if va_h is None else va_h.numpy())
        res[kind] = {"ms": ms, "x_sorted_all_some": ms / base_ms, "parity_ok": bool(ok)}
    # original scores back, then shuffled lefts (any left order is legal: src/granges.rs:935-984)
    for q, ix in enumerate(idxs):
        ix.set_scores(gen_ranges(q, nr_per[q], 14, dev, True)[2])
    perm = torch.cat([int(offs[q]) + torch.randperm(nl_per[q], device=dev) for q in range(len(idxs))])
    ls_s, le_s = ls[perm].contiguous(), le[perm].contiguous()
    call = ctx.prepare_map(idxs, offs, ls_s, le_s, ["mean"], out, ov)
    first_ms = event_ms(call, 1, 0)  # device buffers, AUTO order: the first call takes the lefts as they are and learns from it
    ms = event_ms(call, max(2, steps // 2), 1)
    sc_h = gen_ranges(s, nr_per[s], 14, "cpu", True)[2].numpy()
    ok = check(ls_s, le_s, slice(a, a + nchk), sc_h, None)
    res["shuffled_lefts"] = {"ms": ms, "x_sorted_all_some": ms / base_ms, "parity_ok": bool(ok), "first_call_ms": first_ms,
                             "how": "lefts bucketed by coordinate on the device, rows written back (bucket.cu); order learned from the first call"}
    return res


def run_c3(ctx, idxs, offs, ls, le, nl_per, nr_per, dev, peak):
    """BASELINE config C3 on one GPU: map --func min,max,mean,sum,median over the headline's indexes."""
    from oracle import oracle as orc

    ops = ["min", "max", "mean", "sum", "median"]
    n = len(ls)
    out = torch.empty((n, 5), dtype=torch.float64, device=dev)
    ov = torch.empty((n, 5), dtype=torch.uint8, device=dev)
    call = ctx.prepare_map(idxs, offs, ls, le, ops, out, ov)
    ms = event_ms(call, 3, 1)
    s = CHECK_SEQ
    rs_h, re_h, sc_h = (t.numpy() for t in gen_ranges(s, nr_per[s], 14, "cpu", True))
    a, nchk = int(offs[s]), min(CHECK_LEFTS, nl_per[s])
    hl, he = ls[a:a + nchk].cpu().numpy().astype(np.uint32), le[a:a + nchk].cpu().numpy().astype(np.uint32)
    want, wv = orc.map_joins(orc.Tree(rs_h.astype(np.uint32), re_h.astype(np.uint32)), sc_h, None, hl, he, ops, nthreads=orc.max_threads())
    ok = map_matches_oracle(out[a:a + nchk].cpu().numpy(), ov[a:a + nchk].cpu().numpy(), want, wv, ops)
    nr = sum(nr_per)
    alg = n * 8 + nr * 16 + n * 5 * 8 + n * 5 / 8
    return {"workload": f"map min,max,mean,sum,median {n} x {nr} uniform", "ms": ms, "lefts_per_s": n / ms * 1e3,
            "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg, "parity_ok": bool(ok)}


def run_configs(ctx, dev, peak, scale=1.0):
    """BASELINE configs C2, C4, C5 at their stated sizes on one GPU (own data): ms, fraction of the HBM roofline on SURVEY
    8(d)'s algorithmic bytes, and an oracle spot check on one seqname."""
    import ctypes as C

    from granges_b200 import _capi
    from oracle import oracle as orc

    L_ = _capi.lib()
    res = {}
    # ---- C2: granges filter, 10M x 10M uniform BED3
    n = int(10_000_000 * scale)
    per = split_counts(n, 22)
    idxs, ls, le, chk = [], [], [], None
    for s in range(22):
        a, b = gen_ranges(s, per[s], 13, dev, False)
        rs, re = gen_ranges(s, per[s], 14, dev, False)
        idxs.append(ctx.index_build(rs, re))
        ls.append(a)
        le.append(b)
        if s == CHECK_SEQ:
            chk = (rs.cpu().numpy().astype(np.uint32), re.cpu().numpy().astype(np.uint32), a.cpu().numpy().astype(np.uint32),
                   b.cpu().numpy().astype(np.uint32))
    off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
    ls, le = torch.cat(ls), torch.cat(le)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = event_ms(lambda: ctx.filter_overlaps_batch(idxs, off, ls, le, out=keep), 10, 3)
    a0 = int(off[CHECK_SEQ])
    want = orc.filter_overlaps(orc.Tree(chk[0], chk[1]), chk[2], chk[3])
    ok = bool((keep[a0:a0 + per[CHECK_SEQ]].cpu().numpy() == want).all())
    alg = n * 8 + n * 8 + n
    res["C2"] = {"workload": f"filter {n} x {n} uniform BED3", "ms": ms, "lefts_per_s": n / ms * 1e3, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak,
                 "algorithmic_bytes": alg, "parity_ok": ok}
    for ix in idxs:
        ix.close()
    del idxs, ls, le, keep
    # ---- C4: windows (1 kb / 500 bp) over hg38 x 500M clustered BED5, map mean
    nr = int(500_000_000 * scale)
    nr_per = split_counts(nr, 22)
    idxs, chk, t_build = [], None, 0.0
    for s in range(22):
        rs, re, sc = clustered_rights(s, nr_per[s], 15, dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idxs.append(ctx.index_build(rs, re).set_scores(sc))
        ctx.sync()
        t_build += time.perf_counter() - t0
        if s == CHECK_SEQ:
            chk = (rs.cpu().numpy().astype(np.uint32), re.cpu().numpy().astype(np.uint32), sc.cpu().numpy())
        del rs, re, sc
    w = ctx.windows_dev(HG38, 1000, 500, False)
    off = w.seq_offsets()
    nw = len(w)
    ps, pe = w.dev_ptrs()
    out = torch.empty((nw, 1), dtype=torch.float64, device=dev)
    ov = torch.empty((nw, 1), dtype=torch.uint8, device=dev)
    ops1 = (C.c_int * 1)(4)
    handles = (C.c_void_p * 22)(*[ix.h for ix in idxs])

    def run():
        rc = L_.grb_map_joins_f64_batch(ctx.h, handles, off.ctypes.data_as(_capi.u64p), 22, ps, pe, ops1, 1, out.data_ptr(), ov.data_ptr())
        assert rc == 0, rc

    ms = event_ms(run, 10, 3)
    a0, a1 = int(off[CHECK_SEQ]), int(off[CHECK_SEQ + 1])
    ws = (np.arange(a1 - a0, dtype=np.uint64) * 500).astype(np.uint32)
    we = np.minimum(ws.astype(np.uint64) + 1000, HG38[CHECK_SEQ]).astype(np.uint32)
    want, wv = orc.map_joins(orc.Tree(chk[0], chk[1]), chk[2], None, ws, we, ["mean", "count"], nthreads=orc.max_threads())
    ok = map_matches_oracle(out[a0:a1].cpu().numpy(), ov[a0:a1].cpu().numpy(), want[:, :1], wv[:, :1], ["mean"])
    alg = nr * 16 + nw * 8.125
    res["C4"] = {"workload": f"windows 1 kb / 500 bp ({nw}) x {nr} clustered BED5, map mean", "ms": ms, "lefts_per_s": nw / ms * 1e3,
                 "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg, "parity_ok": bool(ok), "index_build_ms": t_build * 1e3,
                 "max_group_checked": int(want[:, 1].max())}
    for ix in idxs:
        ix.close()
    w.close()
    del idxs, out, ov
    # ---- C5: left_overlaps with full pair materialisation, 50M GTF-like x 200M reads-like, seqname by seqname
    nl, nr = int(50_000_000 * scale), int(200_000_000 * scale)
    nl_per, nr_per = split_counts(nl, 22), split_counts(nr, 22)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    tot_pairs, tot_ms, ok = 0, 0.0, True
    for s in range(22):
        ls, le = gtf_like_lefts(s, nl_per[s], g, dev)
        rs, re = reads_like_rights(s, nr_per[s], g, dev)
        ix = ctx.index_build(rs, re)
        offd = torch.empty(nl_per[s] + 1, dtype=torch.int64, device=dev)
        for rep in range(2):  # the second run reuses the pool's memory: steady state
            ph = C.c_void_p()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = L_.grb_left_overlaps(ctx.h, ix.h, ls.data_ptr(), le.data_ptr(), nl_per[s], offd.data_ptr(), C.byref(ph))
            assert rc == 0, rc
            ctx.sync()
            dt = (time.perf_counter() - t0) * 1e3
            npairs = L_.grb_pairs_len(ph)
            L_.grb_pairs_destroy(ph)
        tot_ms += dt
        tot_pairs += npairs
        if s == 0:
            # spot check: the first 100k lefts of chr1 against all of chr1's rights, offsets and pair indices bit for bit
            nchk = min(100_000, nl_per[s])
            hl, he = ls[:nchk].cpu().numpy().astype(np.uint32), le[:nchk].cpu().numpy().astype(np.uint32)
            tree = orc.Tree(rs.cpu().numpy().astype(np.uint32), re.cpu().numpy().astype(np.uint32))
            woff, widx, _, _ = orc.left_overlaps(tree, hl, he)
            goff, gidx, _, _ = ctx.left_overlaps(ix, ls[:nchk].contiguous(), le[:nchk].contiguous())
            ok = bool((goff == woff).all() and (gidx == widx).all())
            del tree
        ix.close()
        del ls, le, rs, re, offd
    alg = nl * 8 + nr * 12 + (nl + 22) * 8 + tot_pairs * 4
    res["C5"] = {"workload": f"left_overlaps pairs {nl} GTF-like x {nr} reads-like (per seqname, wall clock incl. the pair-count read-back)",
                 "ms": tot_ms, "pairs": tot_pairs, "pairs_per_s": tot_pairs / tot_ms * 1e3, "roofline_frac": alg / (tot_ms * 1e-3) / 1e9 / peak,
                 "algorithmic_bytes": alg, "parity_ok": ok}
    return res


def run_c4_sharded(ctx, dev, rank, world, dist, peak, scale=1.0):
    """BASELINE config C4 (windows 1 kb / 500 bp over hg38 x 500M clustered BED5, map mean) over the ranks of this run:
    the windows are cut into contiguous (seqname, coordinate) units of equal cost, every rank indexes only the rights its
    units can touch and answers its windows; no data moves between ranks.  Device-timed, max over ranks."""
    nr = int(500_000_000 * scale)
    nr_per = split_counts(nr, 22)
    CH = 2048  # planning granularity: chunks of 2048 windows (1.024 Mbp)
    nw = [(HG38[s] + 499) // 500 for s in range(22)]
    rights, chunk_cost, chunk_seq, chunk_w0 = {}, [], [], []
    for s in range(22):
        rs, re, sc = clustered_rights(s, nr_per[s], 15, dev)
        rights[s] = (rs, re, sc)
        edges_w = torch.arange(0, nw[s] + CH, CH, device=dev).clamp(max=nw[s])
        cum = torch.searchsorted(rs.to(torch.int64), edges_w * 500)
        # Cost of a chunk: two searches per window, each a bin-table lookup plus a scan through the bin -- bins are sized for
        # the seqname's MEAN density (about 2 rights), so a window in a region k times denser scans k times further:
        # windows * (1 + occupancy / 8) with occupancy = rights in the chunk / bins in the chunk, bin width = 2 L / n.
        wins = (edges_w[1:] - edges_w[:-1]).to(torch.float64)
        cost = wins + (cum[1:] - cum[:-1]).to(torch.float64) * (HG38[s] / (2000.0 * max(nr_per[s], 1)))
        chunk_cost.append(cost.cpu().numpy())
        chunk_seq += [s] * len(cost)
        chunk_w0 += [int(x) for x in edges_w[:-1].cpu().numpy()]
    cost = np.concatenate(chunk_cost)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    owner = np.minimum((cum[:-1] + cost / 2) / cum[-1] * world, world - 1).astype(np.int64)  # contiguous runs of chunks
    mine = np.nonzero(owner == rank)[0]
    idxs, offs, ls_parts, le_parts, n_right_mine = [], [0], [], [], 0
    if len(mine):
        # runs of my chunks inside one seqname = units
        seqs = np.array(chunk_seq)[mine]
        for s in sorted(set(int(q) for q in seqs)):
            cs = mine[seqs == s]
            w0 = chunk_w0[int(cs[0])]
            w1 = min(chunk_w0[int(cs[-1])] + CH, nw[s])
            L = HG38[s]
            ws = torch.arange(w0, w1, dtype=torch.int64, device=dev) * 500
            we = torch.clamp(ws + 1000, max=L)
            lo, hi = int(ws[0].item()), int(we[-1].item())
            rs, re, sc = rights[s]
            m = (rs < hi) & (re > lo)
            idxs.append(ctx.index_build(rs[m].contiguous(), re[m].contiguous()).set_scores(sc[m].contiguous()))
            n_right_mine += int(m.sum().item())
            ls_parts.append(ws.to(torch.int32))
            le_parts.append(we.to(torch.int32))
            offs.append(offs[-1] + (w1 - w0))
    del rights
    torch.cuda.empty_cache()
    n_mine = offs[-1]
    ms_mine = 0.0
    if n_mine:
        ls, le = torch.cat(ls_parts), torch.cat(le_parts)
        out = torch.empty((n_mine, 1), dtype=torch.float64, device=dev)
        ov = torch.empty((n_mine, 1), dtype=torch.uint8, device=dev)
        call = ctx.prepare_map(idxs, np.array(offs, np.uint64), ls, le, ["mean"], out, ov)
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    if n_mine:
        ms_mine = event_ms(call, 10, 3)
    t = torch.tensor([ms_mine], dtype=torch.float64, device=dev)
    per_rank = [ms_mine]
    if dist is not None:
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = [float(g.item()) for g in gathered]
    agg = torch.tensor([n_mine, n_right_mine], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(agg)
    for ix in idxs:
        ix.close()
    ms = max(per_rank)
    nw_tot = int(agg[0].item())
    alg = nr * 16 + nw_tot * 8.125
    return {"workload": f"windows 1 kb / 500 bp ({nw_tot}) x {nr} clustered BED5, map mean, sharded by (seqname, coordinate) over {world} GPUs",
            "ms": ms, "lefts_per_s": nw_tot / ms * 1e3, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak / world, "algorithmic_bytes": alg,
            "ms_per_rank": per_rank, "rights_indexed_all_ranks": int(agg[1].item())}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every ~5 ms, own thread)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        import threading

        self.sm, self.reason_bits, self.max_mhz, self.err = [], 0, None, None
        self.t_mark = None
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._idx = gpu_index
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        self._ready.wait(timeout=10)

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            uuid = torch.cuda.get_device_properties(self._idx).uuid
            try:
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self._idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._ready.set()
            while not self._stop.is_set():
                t = time.perf_counter()
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                if self.t_mark is not None and t >= self.t_mark:
                    self.sm.append(mhz)
                    self.reason_bits |= bits
                time.sleep(0.002)
            # one more sample right after the region so that a very short region still has one
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self._ready.set()

    def mark(self):
        """Samples from now on belong to the timed region (the thread has been polling since warm-up, so
        NVML start-up cost stays outside the region)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=5)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": len(self.sm)}
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
            out["reasons"] = sorted(n for b, n in self.REASONS.items() if self.reason_bits & b)
        if self.err:
            out["error"] = self.err
        return out


def pick_sample_seqs(nl_per, nr_per, nsample=5):
    """The seqnames whose expected pairs/left are closest to the genome-wide mean (cpu sample)."""
    dens = np.array([nr_per[s] * 1000.0 / HG38[s] for s in range(22)])
    mean = float((dens * np.array(nl_per)).sum() / sum(nl_per))
    order = np.argsort(np.abs(dens - mean), kind="stable")
    return [int(s) for s in order[:nsample]], mean


def cpu_sample_data(seqs, nl_per, nr_per):
    data = []
    for seq in seqs:
        rs, re, sc = (t.numpy() for t in gen_ranges(seq, nr_per[seq], 14, "cpu", True))
        ls, le = (t.numpy() for t in gen_ranges(seq, nl_per[seq], 13, "cpu", False))
        data.append(tuple(a.astype(np.uint32) for a in (rs, re, ls, le)) + (sc,))
    return data


def cpu_reference_run(data, threads, ops=("mean",)):
    """Oracle (coitrees-restated) on whole seqnames: tree build + left_overlaps + map_joins, the
    in-memory work of `granges map` (into_coitrees + left_overlaps + map_joins).  threads == 1 is
    the reference's own execution model (single thread, seqnames one after another); threads > 1
    runs the seqnames concurrently and splits each seqname's lefts over the remaining threads."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as orc

    per_seq = max(1, threads // len(data))

    def one(d):
        rs, re, ls, le, sc = d
        t0 = time.perf_counter()
        tree = orc.Tree(rs, re)
        t1 = time.perf_counter()
        res = orc.map_joins(tree, sc, None, ls, le, list(ops), nthreads=per_seq)
        return t1 - t0, time.perf_counter() - t1, res

    t0 = time.perf_counter()
    if threads == 1:
        parts = [one(d) for d in data]
    else:
        with ThreadPoolExecutor(max_workers=min(threads, len(data))) as ex:
            parts = list(ex.map(one, data))
    total = time.perf_counter() - t0
    return dict(build_s=sum(p[0] for p in parts), map_s=sum(p[1] for p in parts), total_s=total, first=parts[0][2])


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    orc.build()
    nl_per, nr_per = split_counts(args.lefts, 22), split_counts(args.rights, 22)
    seqs, mean_ppl = pick_sample_seqs(nl_per, nr_per)
    data = cpu_sample_data(seqs, nl_per, nr_per)
    n_sample = sum(nl_per[s] for s in seqs)
    threads = orc.max_threads()
    times = []
    for it in range(args.warmup + args.steps):
        r = cpu_reference_run(data, threads, tuple(args.ops.split(",")))
        if it >= args.warmup:
            times.append(r["total_s"])
        last = r
    t = float(np.mean(times))
    value = n_sample / t
    line = {
        "impl": "reference", "metric": "map_mean left ranges/s", "value": value, "unit": "left ranges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "left ranges/s", "cores": threads, "kind": "port",
                         "sample_lefts": n_sample, "sample_rights": sum(nr_per[q] for q in seqs), "sample_seqnames": len(seqs),
                         "extrapolated": "rate measured on the sample (whole seqnames whose pairs/left match the job's mean), quoted as the job's rate; "
                                         "all host threads, although the reference itself is single-threaded",
                         "sample": f"{len(seqs)} whole seqnames ({','.join('chr%d' % (q + 1) for q in seqs)}; {n_sample} lefts x "
                                   f"{sum(nr_per[q] for q in seqs)} rights, ~{mean_ppl:.1f} pairs/left) per step: tree build + "
                                   f"left_overlaps + map_joins, seqnames concurrent, {threads} threads in all "
                                   f"(summed thread time: build {last['build_s']:.2f} s, join+reduce {last['map_s']:.2f} s)"},
        "e2e": {"value": value, "unit": "left ranges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"granges map --func mean: {args.lefts} left x {args.rights} right synthetic hg38 ranges (22 autosomes, "
                        "len U{1..999}, sorted BED3 x BED5)",
            "lefts": args.lefts, "rights": args.rights, "ops": ["mean"], "l2": "inputs_larger_than_l2",
            "sharding": "seqname/coordinate units, no collective"}


def bind_to_gpu_numa_node(device_index):
    """Run this process on the CPUs next to its GPU (what `numactl` would do): the pinned host buffers of the e2e leg
    are then first-touched on the GPU's NUMA node instead of wherever the scheduler happened to start the process."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        uuid = torch.cuda.get_device_properties(device_index).uuid
        try:
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(device_index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {w * 64 + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if len(cpus) >= 8:  # the host path runs five threads; never squeeze them onto a handful of cores
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lefts", type=int, default=100_000_000)
    ap.add_argument("--rights", type=int, default=100_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `variants` and `configs` legs (N = 1 only)")
    ap.add_argument("--ops", default="mean")
    ap.add_argument("--valid-bits", action="store_true", help="validity as one bit per result (grb_map_joins_f64_batch_bits); measured slower "
                                                              "than bytes on this kernel (profiles/r2/ab_pipe_kernel.txt), so bytes are the default")
    ap.add_argument("--valid-bytes", action="store_true", help="(default) validity as one byte per result")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: granges_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # NCCL may print its version banner on stdout when the communicator comes up; the contract is ONE JSON
        # line on stdout, so stdout points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
            # a host-only group: ranks that wait while rank 0 drives every GPU from one process must not wait inside an NCCL
            # kernel (it would share their GPU's time slices with rank 0's work there)
            cpu_group = dist.new_group(backend="gloo")
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    import granges_b200 as gb

    ops = args.ops.split(",")
    k = len(ops)
    nl_per, nr_per = split_counts(args.lefts, 22), split_counts(args.rights, 22)
    left_off = np.concatenate([[0], np.cumsum(nl_per)]).astype(np.uint64)

    # ---- shard plan (identical on every rank; bounds come from the data below)
    from granges_b200 import sharding

    # cost of a seqname = lefts + rights, plus its expected pairs when a reduction has to visit the members
    # (min / max / median / weighted mean): pairs = NL * NR * (mean left length + mean right length) / seqname length
    # (prefix-sum path: a tile of 1024 lefts also stages the rights its windows reach back over -- ~1000 bp worth, i.e.
    # 1000 * rights / seqname length per tile: 2 % of a tile's rights on chr1, 10 % on chr21 -- measured as the spread of the
    # per-rank step times at N = 8 with the plain lefts + rights cost)
    cost_per = [int(nr_per[s] + (nl_per[s] / 1024.0) * 1000.0 * nr_per[s] / HG38[s]) for s in range(22)]
    # (round 1 added PAIR_WEIGHT * expected pairs for min / max / median: the member sweep's cost grew with the group sizes.  The
    # wavelet runs, the stabbing tables and the closed forms of round 2 cost the same per left whatever the coverage, so dense
    # lefts over dense rights -- this job -- are planned like the prefix-sum path; the weight remains for sparse lefts, where
    # the sweep still answers: GRB_BENCH_PAIR_WEIGHT)
    pair_weight = float(os.environ.get("GRB_BENCH_PAIR_WEIGHT", "0"))
    if pair_weight > 0 and any(o in ("min", "max", "median", "weighted-mean") for o in ops):
        cost_per = [int(cost_per[s] + pair_weight * nl_per[s] * nr_per[s] * 1000.0 / HG38[s]) for s in range(22)]
    # (development knob: time the share rank R of a W-rank run would get, on one GPU, without launching W processes)
    plan_world, plan_rank = int(os.environ.get("GRB_BENCH_AS_WORLD", world)), int(os.environ.get("GRB_BENCH_AS_RANK", rank))
    plan = sharding.plan(left_off, cost_per, plan_world)
    mine = sharding.units_of(plan, plan_rank)

    # a non-default stream: stream-ordered allocation on the legacy default stream is several times slower
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = gb.Context(local_rank, stream=stream)
    # ---- synthetic data, generated on the device, only what this rank's units touch
    idxs, offs, ls_parts, le_parts = [], [0], [], []
    n_right_mine = 0
    t_build = 0.0
    for p in mine:
        s = p["seq"]
        a, b = p["left_begin"] - int(left_off[s]), p["left_end"] - int(left_off[s])
        ls_s, le_s = gen_ranges(s, nl_per[s], 13, dev, False)
        ls_s, le_s = ls_s[a:b].contiguous(), le_s[a:b].contiguous()
        rs, re, sc = gen_ranges(s, nr_per[s], 14, dev, True)
        if (a, b) != (0, nl_per[s]):
            # a coordinate slice of the seqname: keep the rights that can overlap this run of lefts
            lo, hi = sharding.unit_bounds(ls_s, le_s)
            m = sharding.unit_right_mask(rs, re, lo, hi)
            rs, re, sc = rs[m].contiguous(), re[m].contiguous(), sc[m].contiguous()
            # re-base the unit to its own origin: overlaps only depend on coordinate differences, and the index sizes its
            # coordinate bins for [0, max end) -- a unit from the far end of a chromosome would get bins several rights wide
            base = sharding.unit_origin(rs, lo)
            rs, re, ls_s, le_s = rs - base, re - base, ls_s - base, le_s - base
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix = ctx.index_build(rs, re).set_scores(sc)
        ctx.sync()
        t_build += time.perf_counter() - t0
        idxs.append(ix)
        n_right_mine += len(rs)
        offs.append(offs[-1] + (b - a))
        ls_parts.append(ls_s)
        le_parts.append(le_s)
        del rs, re, sc
    ls = torch.cat(ls_parts) if ls_parts else torch.zeros(0, dtype=torch.int32, device=dev)
    le = torch.cat(le_parts) if le_parts else torch.zeros(0, dtype=torch.int32, device=dev)
    del ls_parts, le_parts
    nl_mine = int(offs[-1])
    offs = np.array(offs, dtype=np.uint64)
    out = torch.empty((nl_mine, k), dtype=torch.float64, device=dev)
    bits = args.valid_bits
    ov = torch.empty(((nl_mine * k + 31) // 32,), dtype=torch.int32, device=dev) if bits else torch.empty((nl_mine, k), dtype=torch.uint8, device=dev)

    call = ctx.prepare_map(idxs, offs, ls, le, ops, out, ov, valid_bits=bits) if nl_mine else None

    def step():
        if call is not None:
            call()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark()
    t_all0.record()
    for e0, e1 in evs:
        e0.record()
        step()
        e1.record()
    t_all1.record()
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = t_all0.elapsed_time(t_all1)
    per_step = [a.elapsed_time(b) for a, b in evs]
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    per_rank_ms = [total_ms / args.steps]
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor([total_ms / args.steps], dtype=torch.float64, device=dev))
        per_rank_ms = [float(g.item()) for g in gathered]
    total_ms_max = float(tmax.item())
    ms_per_step = total_ms_max / args.steps

    # ---- results: pairs (for pairs/s) and a checksum
    pairs_mine = 0
    if nl_mine:
        cnt = ctx.count_overlaps_batch(idxs, offs, ls, le)
        ctx.sync()
        pairs_mine = int(cnt.to(torch.int64).sum().item())
        del cnt
    agg = torch.tensor([pairs_mine, nl_mine, n_right_mine], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(agg)
    pairs_total, nl_total = int(agg[0].item()), int(agg[1].item())

    # ---- index build once more on the warm memory pool (index_build_ms_rank0 above is the first build of the process: it
    # pays for the pool's growth): the same units, one by one with a synchronisation each, and closed again
    t_build_warm = 0.0
    if world == 1 and not args.no_configs:
        for p in mine:
            s = p["seq"]
            if (p["left_begin"] - int(left_off[s]), p["left_end"] - int(left_off[s])) != (0, nl_per[s]):
                t_build_warm = None  # (coordinate slices: not re-derived here)
                break
            rs, re, sc = gen_ranges(s, nr_per[s], 14, dev, True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ixw = ctx.index_build(rs, re).set_scores(sc)
            ctx.sync()
            t_build_warm += time.perf_counter() - t0
            ixw.close()
            del rs, re, sc

    # ... and all of them through grb_index_build_batch (four host threads with private streams), second of two passes
    t_build_batch = None
    if world == 1 and not args.no_configs and t_build_warm:
        rights_all = [gen_ranges(p["seq"], nr_per[p["seq"]], 14, dev, True) for p in mine]
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ixs = ctx.index_build_batch(rights_all)
            ctx.sync()
            t_build_batch = time.perf_counter() - t0
            for ixw in ixs:
                ixw.close()
        del rights_all, ixs

    # keep one seqname's device-timed results for the parity check against the oracle below
    chk_seq = pick_sample_seqs(nl_per, nr_per)[0][0]
    chk_slice = None
    if world == 1 and rank == 0:
        a, b = int(left_off[chk_seq]), int(left_off[chk_seq + 1])
        if bits:
            w0, w1 = (a * k) // 32, (b * k + 31) // 32
            flat = np.unpackbits(ov[w0:w1].cpu().numpy().view(np.uint8), bitorder="little")[a * k - w0 * 32: b * k - w0 * 32]
            chk_slice = (out[a:b].cpu().numpy(), flat.reshape(b - a, k))
        else:
            chk_slice = (out[a:b].cpu().numpy(), ov[a:b].cpu().numpy())

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- the other input classes of the headline job, and config C3, on the same indexes (N = 1)
    variants = configs = None
    if world == 1 and not args.no_configs and ops == ["mean"] and nl_mine:
        call = None
        variants = run_variants(ctx, idxs, offs, ls, le, nl_per, nr_per, dev, float(np.mean(per_step)), args.steps, args.warmup)
        configs = {"C3": run_c3(ctx, idxs, offs, ls, le, nl_per, nr_per, dev, peak)}

    # ---- e2e: the whole job from HOST buffers through ONE call of the C ABI -- grb_map_joins_whole_f64 at N = 1,
    # grb_map_joins_whole_f64_mgpu over the N GPUs of the box at N > 1 (one process: rank 0 makes the call from the host
    # arrays of the WHOLE job, the gather into one output array is inside the timed region; the other ranks free their
    # GPU memory and wait).  Measured for page-locked and for pageable host buffers.
    e2e = None
    if not args.no_e2e:
        call = None
        for ix in idxs:
            ix.close()
        idxs.clear()
        del out, ov, ls, le
        torch.cuda.empty_cache()
        barrier()

        def host_barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier(group=cpu_group)

        host_barrier()
        if rank == 0:
            host = []
            for s in range(22):
                rs, re, sc = gen_ranges(s, nr_per[s], 14, dev, True)
                a, b = gen_ranges(s, nl_per[s], 13, dev, False)
                host.append(tuple(t.cpu().numpy() for t in (rs, re, sc, a, b)))
                del rs, re, sc, a, b
            torch.cuda.empty_cache()
            p_ls = np.concatenate([h[3] for h in host]).view(np.uint32)
            p_le = np.concatenate([h[4] for h in host]).view(np.uint32)
            p_rights = [(h[0].view(np.uint32), h[1].view(np.uint32), h[2]) for h in host]
            p_out, p_ov = np.empty((args.lefts, k), np.float64), np.empty((args.lefts, k), np.uint8)
            pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
            h_ls, h_le = pin(p_ls), pin(p_le)
            h_rights = [tuple(pin(x) for x in r) for r in p_rights]
            h_out, h_ov = pin(p_out), pin(p_ov)
            devices = list(range(world))

            def e2e_step(rights, l_s, l_e, o, v):
                if world == 1:
                    ctx.map_whole(rights, left_off, l_s, l_e, ops, out=o, out_valid=v)
                else:
                    gb.map_whole_mgpu(devices, rights, left_off, l_s, l_e, ops, out=o, out_valid=v)
                return float(o[0, 0])  # the result is on the host

            def timed(args_):
                e2e_step(*args_)
                n = max(2, min(args.steps, 5))
                t0 = time.perf_counter()
                for _ in range(n):
                    e2e_step(*args_)
                return (time.perf_counter() - t0) / n, n

            e2e_s, e2e_steps = timed((h_rights, h_ls, h_le, h_out, h_ov))
            pg_s, _ = timed((p_rights, p_ls, p_le, p_out, p_ov))
            same = bool((p_ov == h_ov).all() and (p_out[p_ov == 1] == h_out[h_ov == 1]).all())
            e2e = {"value": args.lefts / e2e_s, "unit": "left ranges/s", "h2d_bytes_per_step": args.rights * 16 + args.lefts * 8,
                   "d2h_bytes_per_step": args.lefts * k * 9, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                   "pageable": {"value": args.lefts / pg_s, "ms_per_step": pg_s * 1e3, "equals_pinned": same,
                                "how": "the same call on ordinary (pageable) numpy arrays: the library stages them through its own page-locked rings"},
                   "includes": ("grb_map_joins_whole_f64" if world == 1 else f"grb_map_joins_whole_f64_mgpu over {world} GPUs (one process, one host "
                                "thread + context per GPU, contiguous runs of seqnames, every GPU writes its rows of the one output array)") +
                               " = `granges map` in one call: H2D of rights + scores + lefts, index builds, map pass, D2H of the results, pipelined "
                               "per seqname; page-locked host buffers; wall clock of the calling process"}
            del h_rights, h_ls, h_le, p_rights, p_ls, p_le, host
        host_barrier()

    c4_sharded = None
    if world > 1 and not args.no_configs and ops == ["mean"]:
        c4_sharded = run_c4_sharded(ctx, dev, rank, world, dist, peak)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = args.lefts / (ms_per_step * 1e-3)
    line = {
        "metric": "map_mean left ranges/s", "value": value, "unit": "left ranges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "pairs_per_s": pairs_total / (ms_per_step * 1e-3), "overlap_pairs": pairs_total, "lefts_processed": nl_total,
        "index_build_ms_rank0": t_build * 1e3, "index_build_warm_ms_rank0": (t_build_warm * 1e3 if t_build_warm else None),
        "index_build_batch_ms_rank0": (t_build_batch * 1e3 if t_build_batch else None),
        "gpu_launches": int(launches), "clocks": clocks,
        "step_ms_rank0": {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step)},
        "units_rank0": len(mine), "cpus_bound_rank0": numa, "step_ms_per_rank": per_rank_ms,
    }
    if ops != ["mean"]:
        line["metric"] = "map_%s left ranges/s" % "_".join(ops)
        line["config"]["ops"] = ops

    # ---- roofline of the dominant kernel (k_map_pipe<MEAN>: the only kernel of a map_mean step)
    nr_tot = int(agg[2].item())
    alg_bytes = nl_total * 8 + nr_tot * 16 + nl_total * k * 8 + nl_total * k / 8  # B_map
    kernel_ms = float(np.mean(per_step))  # rank 0's launch duration (one k_map_pipe launch per step when ops = mean)
    alg_rank0 = nl_mine * 8 + n_right_mine * 16 + nl_mine * k * 8 + nl_mine * k / 8
    achieved = alg_rank0 / (kernel_ms * 1e-3) / 1e9
    line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "peak_source": peak_src, "kernel": "k_map_pipe<MEAN>" if ops == ["mean"] else "k_tile/k_sweep", "kernel_ms": kernel_ms,
                        "algorithmic_bytes_per_launch": alg_rank0, "algorithmic_bytes_whole_job": alg_bytes}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            if tj.get("lefts") == nl_mine:
                line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
        except Exception:
            pass

    line["e2e"] = e2e
    if variants is not None:
        line["variants"] = variants
    if configs is not None:
        torch.cuda.empty_cache()
        configs.update(run_configs(ctx, dev, peak))
        line["configs"] = configs
    if c4_sharded is not None:
        line["configs"] = {"C4": c4_sharded}

    # ---- cpu baseline: the oracle on a bounded sample (one seqname), 1 thread like the reference
    if world == 1 and not args.no_cpu:
        seqs, mean_ppl = pick_sample_seqs(nl_per, nr_per)
        data = cpu_sample_data(seqs, nl_per, nr_per)
        n_sample = sum(nl_per[q] for q in seqs)
        r = cpu_reference_run(data, 1, tuple(ops))
        want, wv = r["first"]
        got, gv = chk_slice
        okv = bool((gv == wv).all())
        m = wv == 1
        rel = float(np.max(np.abs(got[m] - want[m]) / np.maximum(np.abs(want[m]), 1e-300))) if m.any() else 0.0
        line["parity_check"] = {"seqname": f"chr{chk_seq + 1}", "lefts": int(len(want)), "validity_equal": okv, "max_rel_err": rel,
                                "tolerance": 1e-12, "ok": bool(okv and rel <= 1e-12)}
        line["cpu_baseline"] = {"value": n_sample / r["total_s"], "unit": "left ranges/s", "cores": 1, "kind": "port",
                                "sample_lefts": n_sample, "sample_rights": sum(nr_per[q] for q in seqs), "sample_seqnames": len(seqs),
                                "sample": f"{len(seqs)} whole seqnames ({','.join('chr%d' % (q + 1) for q in seqs)}; {n_sample} lefts x "
                                          f"{sum(nr_per[q] for q in seqs)} rights, ~{mean_ppl:.1f} pairs/left): tree build "
                                          f"{r['build_s']:.2f} s + left_overlaps + map_joins {r['map_s']:.2f} s, single thread "
                                          f"like the reference (host has {os.cpu_count()} cores)"}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
