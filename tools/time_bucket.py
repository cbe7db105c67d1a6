"""Per-kernel view of the unsorted-lefts path at bench size (run under ncu --metrics gpu__time_duration.sum)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import granges_b200 as gb
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = gb.Context(0)
dev = torch.device("cuda", 0)
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
ctx.set_left_order("unsorted")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.map_joins_batch(idxs, off, ls, le, ["mean"], out=out, out_valid=ov)
    ctx.sync(); torch.cuda.synchronize()
    print("unsorted", rep, "ms", round((time.perf_counter() - t0) * 1e3, 3), flush=True)
ctx.set_left_order("auto")
call = ctx.prepare_map(idxs, off, ls, le, ["mean"], out, ov)
for rep in range(4):
    n0 = ctx.launch_count
    torch.cuda.synchronize(); t0 = time.perf_counter()
    call()
    ctx.sync(); torch.cuda.synchronize()
    print("auto (prepared)", rep, "ms", round((time.perf_counter() - t0) * 1e3, 3), "launches", ctx.launch_count - n0, flush=True)
