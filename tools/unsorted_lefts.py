"""How much does left order matter?  The same 100M x 100M map_mean with the lefts of every seqname shuffled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, granges_b200 as gb
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = gb.Context(0, stream=st)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
per = bench.split_counts(n, 22)
idxs, ls, le, lsh, leh = [], [], [], [], []
g = torch.Generator(device=dev); g.manual_seed(1)
for s in range(22):
    a, b = bench.gen_ranges(s, per[s], 13, dev, False)
    rs, re, sc = bench.gen_ranges(s, per[s], 14, dev, True)
    idxs.append(ctx.index_build(rs, re).set_scores(sc))
    p = torch.randperm(per[s], generator=g, device=dev)
    ls.append(a); le.append(b); lsh.append(a[p]); leh.append(b[p])
off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
for name, L, E in (("sorted", torch.cat(ls), torch.cat(le)), ("shuffled", torch.cat(lsh), torch.cat(leh))):
    out = torch.empty((n, 1), dtype=torch.float64, device=dev); ov = torch.empty((n, 1), dtype=torch.uint8, device=dev)
    call = ctx.prepare_map(idxs, off, L, E, ["mean"], out, ov)
    call(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); call(); call(); e1.record(); ctx.sync(); torch.cuda.synchronize()
    print(name, "ms per pass: %.3f" % (e0.elapsed_time(e1) / 3), "checksum %.6f" % float(out[ov.bool()].sum()))
