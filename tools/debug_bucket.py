"""Debug: launch counts and timings of the left-order paths."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import granges_b200 as gb
from tests.randcases import random_case
import bench

ctx = gb.Context(0)
c = random_case(801, 150_000, 120_000, span=3_000_000, maxlen=400)
ix = ctx.index_build(c["rs"], c["re"]).set_scores(c["scores"])
off = np.array([0, len(c["ls"])], np.uint64)
for order in ("auto", "sorted", "unsorted", "auto"):
    ctx.set_left_order(order)
    n0 = ctx.launch_count
    t0 = time.perf_counter()
    ctx.map_joins_batch([ix], off, c["ls"], c["le"], ["mean"])
    print(order, "launches", ctx.launch_count - n0, "ms", (time.perf_counter() - t0) * 1e3)
# big device case
dev = torch.device("cuda", 0)
n = 20_000_000
per = bench.split_counts(n, 22)
idxs, ls, le = [], [], []
for s in range(22):
    a, b = bench.gen_ranges(s, per[s], 13, dev, False)
    rs, re, sc = bench.gen_ranges(s, per[s], 14, dev, True)
    idxs.append(ctx.index_build(rs, re).set_scores(sc))
    p = torch.randperm(per[s], device=dev)
    ls.append(a[p]); le.append(b[p])
off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
ls, le = torch.cat(ls), torch.cat(le)
out = torch.empty((n, 1), dtype=torch.float64, device=dev); ov = torch.empty((n, 1), dtype=torch.uint8, device=dev)
for order in ("sorted", "unsorted", "auto", "auto", "auto"):
    ctx.set_left_order(order) if order != "auto" or True else None
    for rep in range(3):
        n0 = ctx.launch_count
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.map_joins_batch(idxs, off, ls, le, ["mean"], out=out, out_valid=ov)
        ctx.sync(); torch.cuda.synchronize()
        print(order, rep, "launches", ctx.launch_count - n0, "ms", round((time.perf_counter() - t0) * 1e3, 3))
