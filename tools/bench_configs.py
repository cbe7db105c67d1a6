"""BASELINE.json configs C1..C5 at (or near) full size on ONE B200: parity-test shapes, timed for the record.
Not the bench line (bench.py is); results go to profiles/.  Each config prints one JSON line.
usage: python tools/bench_configs.py [c1 c2 c3 c4 c5] [--scale F]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import granges_b200 as gb

HG38 = bench.HG38
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = gb.Context(0, stream=stream)


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        ctx.sync()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), r


def uniform_sets(nl, nr, with_scores=True):
    nl_per, nr_per = bench.split_counts(nl, 22), bench.split_counts(nr, 22)
    idxs, ls, le, t_build = [], [], [], 0.0
    for s in range(22):
        a, b = bench.gen_ranges(s, nl_per[s], 13, dev, False)
        ls.append(a)
        le.append(b)
        if with_scores:
            rs, re, sc = bench.gen_ranges(s, nr_per[s], 14, dev, True)
        else:
            rs, re = bench.gen_ranges(s, nr_per[s], 14, dev, False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix = ctx.index_build(rs, re)
        if with_scores:
            ix.set_scores(sc)
        ctx.sync()
        t_build += time.perf_counter() - t0
        idxs.append(ix)
    off = np.concatenate([[0], np.cumsum(nl_per)]).astype(np.uint64)
    return idxs, off, torch.cat(ls), torch.cat(le), t_build


def report(name, **kw):
    print(json.dumps(dict(config=name, **kw)), flush=True)


def c1():
    idxs, off, ls, le, tb = uniform_sets(100_000, 100_000)
    ms, (out, ov) = timed(lambda: ctx.map_joins_batch(idxs, off, ls, le, ["mean"]))
    report("C1 map mean 100k x 100k", ms=ms, lefts_per_s=100_000 / ms * 1e3, index_build_ms=tb * 1e3, some=int(ov.sum().item()))


def c2(scale):
    n = int(10_000_000 * scale)
    idxs, off, ls, le, tb = uniform_sets(n, n, with_scores=False)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    ms, (_, nk) = timed(lambda: ctx.filter_overlaps_batch(idxs, off, ls, le, out=keep))
    report("C2 filter %d x %d" % (n, n), ms=ms, lefts_per_s=n / ms * 1e3, kept=int(nk), index_build_ms=tb * 1e3)


def c3(scale):
    n = int(100_000_000 * scale)
    idxs, off, ls, le, tb = uniform_sets(n, n)
    ops = ["min", "max", "mean", "sum", "median"]
    out = torch.empty((n, 5), dtype=torch.float64, device=dev)
    ov = torch.empty((n, 5), dtype=torch.uint8, device=dev)
    ms, _ = timed(lambda: ctx.map_joins_batch(idxs, off, ls, le, ops, out=out, out_valid=ov), reps=2)
    cnt = ctx.count_overlaps_batch(idxs, off, ls, le)
    pairs = int(cnt.to(torch.int64).sum().item())
    report("C3 map min,max,mean,sum,median %d x %d (1 GPU)" % (n, n), ms=ms, lefts_per_s=n / ms * 1e3, pairs=pairs,
           pairs_per_s=pairs / ms * 1e3, index_build_ms=tb * 1e3)


def c3ops(scale):
    """C3's member sweep by reduction subset: where the time goes."""
    n = int(100_000_000 * scale)
    idxs, off, ls, le, tb = uniform_sets(n, n)
    for ops in (["min"], ["min", "max"], ["median"], ["min", "max", "median"], ["min", "max", "mean", "sum", "median"], ["overlap-bp"], ["weighted-mean"]):
        out = torch.empty((n, len(ops)), dtype=torch.float64, device=dev)
        ov = torch.empty((n, len(ops)), dtype=torch.uint8, device=dev)
        ms, _ = timed(lambda: ctx.map_joins_batch(idxs, off, ls, le, ops, out=out, out_valid=ov), reps=2)
        report("C3 subset " + ",".join(ops) + " %d x %d" % (n, n), ms=ms, lefts_per_s=n / ms * 1e3)


def clustered_rights(seq, n, seed):
    """80 % of the ranges in 5 % of the 1 Mb bins (log-normal bin weights, sigma 1.5), the rest uniform."""
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


def c4(scale):
    nr = int(500_000_000 * scale)
    nr_per = bench.split_counts(nr, 22)
    idxs, tb = [], 0.0
    for s in range(22):
        rs, re, sc = clustered_rights(s, nr_per[s], 15)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idxs.append(ctx.index_build(rs, re).set_scores(sc))
        ctx.sync()
        tb += time.perf_counter() - t0
        del rs, re, sc
    t0 = time.perf_counter()
    w = ctx.windows_dev(HG38, 1000, 500, False)
    ctx.sync()
    t_win = time.perf_counter() - t0
    off = w.seq_offsets()
    n = len(w)
    ps, pe = w.dev_ptrs()
    import ctypes as C

    from granges_b200 import _capi

    out = torch.empty((n, 1), dtype=torch.float64, device=dev)
    ov = torch.empty((n, 1), dtype=torch.uint8, device=dev)
    ops = (C.c_int * 1)(4)
    handles = (C.c_void_p * 22)(*[ix.h for ix in idxs])

    def run():
        rc = _capi.lib().grb_map_joins_f64_batch(ctx.h, handles, off.ctypes.data_as(_capi.u64p), 22, ps, pe, ops, 1, out.data_ptr(),
                                                 ov.data_ptr())
        assert rc == 0

    ms, _ = timed(run)
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    rc = _capi.lib().grb_count_overlaps_batch(ctx.h, handles, off.ctypes.data_as(_capi.u64p), 22, ps, pe, cnt.data_ptr())
    assert rc == 0
    ctx.sync()
    pairs = int(cnt.to(torch.int64).sum().item())
    report("C4 windows(1kb/500bp) x %d clustered BED5, map mean" % nr, windows=n, windows_ms=t_win * 1e3, ms=ms, lefts_per_s=n / ms * 1e3,
           pairs=pairs, max_group=int(cnt.max().item()), pairs_per_s=pairs / ms * 1e3, index_build_ms=tb * 1e3)


def c4ops(scale):
    """C4's right set (500M clustered BED5) under the member reductions: skewed groups (hundreds to tens of thousands of
    members per window) through k_minmax, k_sweep3 with whole-warp groups and the radix-select path for the heaviest."""
    nr = int(500_000_000 * scale)
    nr_per = bench.split_counts(nr, 22)
    idxs = []
    for s in range(22):
        rs, re, sc = clustered_rights(s, nr_per[s], 15)
        idxs.append(ctx.index_build(rs, re).set_scores(sc))
        del rs, re, sc
    w = ctx.windows_dev(HG38, 1000, 500, False)
    off = w.seq_offsets()
    n = len(w)
    ps, pe = w.dev_ptrs()
    import ctypes as C

    from granges_b200 import _capi

    for ops_names in (["min", "max"], ["median"], ["overlap-bp"], ["weighted-mean"]):
        k = len(ops_names)
        out = torch.empty((n, k), dtype=torch.float64, device=dev)
        ov = torch.empty((n, k), dtype=torch.uint8, device=dev)
        ops = (C.c_int * k)(*[gb.OPS[o] for o in ops_names])
        handles = (C.c_void_p * 22)(*[ix.h for ix in idxs])

        def run():
            rc = _capi.lib().grb_map_joins_f64_batch(ctx.h, handles, off.ctypes.data_as(_capi.u64p), 22, ps, pe, ops, k, out.data_ptr(),
                                                     ov.data_ptr())
            assert rc == 0, ctx._check(rc)

        t0 = time.perf_counter()
        run()
        ctx.sync()
        first = (time.perf_counter() - t0) * 1e3
        ms, _ = timed(run, reps=2, warm=0)
        report("C4 rights (%d clustered) x %d windows, %s" % (nr, n, ",".join(ops_names)), ms=ms, first_call_ms_incl_tables=first,
               lefts_per_s=n / ms * 1e3, checksum=float(out[ov.bool()].sum().item()))


def c5(scale):
    nl, nr = int(50_000_000 * scale), int(200_000_000 * scale)
    nl_per, nr_per = bench.split_counts(nl, 22), bench.split_counts(nr, 22)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    tot_pairs, tot_ms, tb = 0, 0.0, 0.0
    for s in range(22):
        L = HG38[s]
        # GTF-like lefts: log-normal(median 2 kb, sigma 1.0) clipped to [50, 2e6]
        ln = torch.exp(np.log(2000.0) + torch.randn(nl_per[s], generator=g, device=dev, dtype=torch.float64)).clamp(50, 2e6).to(torch.int64)
        st = (torch.rand(nl_per[s], generator=g, device=dev, dtype=torch.float64) * (L - ln).to(torch.float64)).to(torch.int64)
        key, _ = torch.sort(st * 4294967296 + (st + ln))
        ls, le = (key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32)
        # reads-like rights: len U{75..151}
        rl = torch.randint(75, 152, (nr_per[s],), generator=g, device=dev)
        rst = (torch.rand(nr_per[s], generator=g, device=dev, dtype=torch.float64) * (L - 151)).to(torch.int64)
        key, _ = torch.sort(rst * 4294967296 + (rst + rl))
        rs, re = (key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix = ctx.index_build(rs, re)
        ctx.sync()
        tb += time.perf_counter() - t0
        import ctypes as C

        from granges_b200 import _capi

        off = torch.empty(nl_per[s] + 1, dtype=torch.int64, device=dev)
        for rep in range(2):  # the second run reuses the pool's memory: steady state
            ph = C.c_void_p()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = _capi.lib().grb_left_overlaps(ctx.h, ix.h, ls.data_ptr(), le.data_ptr(), nl_per[s], off.data_ptr(), C.byref(ph))
            assert rc == 0, ctx._check(rc)
            ctx.sync()
            dt = (time.perf_counter() - t0) * 1e3
            npairs = _capi.lib().grb_pairs_len(ph)
            _capi.lib().grb_pairs_destroy(ph)
        tot_ms += dt
        tot_pairs += npairs
        ix.close()
    report("C5 left_overlaps pairs %d x %d (per seqname, wall clock of the second run incl. allocation from the pool and the P read-back)" % (nl, nr), ms=tot_ms, pairs=tot_pairs,
           pairs_per_s=tot_pairs / tot_ms * 1e3, lefts_per_s=nl / tot_ms * 1e3, index_build_ms=tb * 1e3)


def c6(scale):
    """granges merge: sorted BED5 records, distance 0, sum of the merged scores; device-resident input."""
    n = int(100_000_000 * scale)
    per = bench.split_counts(n, 22)
    g = torch.Generator(device=dev)
    g.manual_seed(6)
    ss, ee = [], []
    for s in range(22):
        L = HG38[s]
        st, _ = torch.sort((torch.rand(per[s], generator=g, device=dev, dtype=torch.float64) * (L - 40)).to(torch.int64))
        ln = torch.randint(1, 40, (per[s],), generator=g, device=dev)
        ss.append(st.to(torch.int32))
        ee.append((st + ln).to(torch.int32))
    st, en = torch.cat(ss), torch.cat(ee)
    sc = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    ro = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
    import ctypes as C

    from granges_b200 import _capi

    L_ = _capi.lib()
    ops = (C.c_int * 1)(0)
    res = {}

    def run():
        h = C.c_void_p()
        t0 = time.perf_counter()
        rc = L_.grb_merge_ranges(ctx.h, st.data_ptr(), en.data_ptr(), ro.ctypes.data_as(_capi.u64p), 22, 0, C.byref(h))
        assert rc == 0, ctx._check(rc)
        ctx.sync()
        t1 = time.perf_counter()
        m = L_.grb_merged_len(h)
        out = torch.empty(m, dtype=torch.float64, device=dev)
        ov = torch.empty(m, dtype=torch.uint8, device=dev)
        rc = L_.grb_merged_map_f64(ctx.h, h, sc.data_ptr(), None, ops, 1, out.data_ptr(), ov.data_ptr())
        assert rc == 0, ctx._check(rc)
        ctx.sync()
        t2 = time.perf_counter()
        L_.grb_merged_destroy(h)
        res.update(m=m, merge_ms=(t1 - t0) * 1e3, map_ms=(t2 - t1) * 1e3, total=float(out.sum().item()))

    run()
    run()
    report("C6 merge %d sorted BED5 records, distance 0, sum (wall clock of the second run)" % n, merged=res["m"], merge_ms=res["merge_ms"],
           reduce_ms=res["map_ms"], records_per_s=n / (res["merge_ms"] + res["map_ms"]) * 1e3, score_total=res["total"],
           score_total_check=float(sc.sum().item()))


if __name__ == "__main__":
    argv = sys.argv[1:]
    scale = 1.0
    if "--scale" in argv:
        i = argv.index("--scale")
        scale = float(argv[i + 1])
        del argv[i:i + 2]
    args = argv
    which = args or ["c1", "c2", "c3", "c4", "c5", "c6"]
    for w in which:
        torch.cuda.empty_cache()
        {"c1": c1, "c2": lambda: c2(scale), "c3": lambda: c3(scale), "c4": lambda: c4(scale), "c5": lambda: c5(scale), "c3ops": lambda: c3ops(scale), "c6": lambda: c6(scale), "c4ops": lambda: c4ops(scale)}[w]()
