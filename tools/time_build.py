"""Index build of the bench's right set (100M sorted BED5 rights, 22 seqnames), device-resident inputs:
wall time per pass, one seqname after another and through grb_index_build_batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import granges_b200 as gb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = gb.Context(0, stream=stream)
per = bench.split_counts(n, 22)
rights = [bench.gen_ranges(s, per[s], 14, dev, True) for s in range(22)]
torch.cuda.synchronize()
for rep in range(reps):
    t0 = time.perf_counter()
    idxs = [ctx.index_build(rs, re).set_scores(sc) for rs, re, sc in rights]
    ctx.sync()
    t1 = time.perf_counter()
    for ix in idxs:
        ix.close()
    print(f"one by one  rep {rep}: {1e3 * (t1 - t0):8.2f} ms", flush=True)
for rep in range(reps):
    t0 = time.perf_counter()
    idxs = ctx.index_build_batch(rights)
    ctx.sync()
    t1 = time.perf_counter()
    for ix in idxs:
        ix.close()
    print(f"batch       rep {rep}: {1e3 * (t1 - t0):8.2f} ms", flush=True)
