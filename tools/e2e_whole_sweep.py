"""grb_map_joins_whole_f64 at full size under different chunk / worker counts (env knobs read per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, granges_b200 as gb

n = 100_000_000
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = gb.Context(0, stream=st)
per = bench.split_counts(n, 22)
rights, ls, le = [], [], []
for s in range(22):
    a, b = bench.gen_ranges(s, per[s], 13, dev, False)
    rs, re, sc = bench.gen_ranges(s, per[s], 14, dev, True)
    rights.append(tuple(t.cpu().pin_memory() for t in (rs, re, sc)))
    ls.append(a.cpu()); le.append(b.cpu())
h_ls = torch.cat(ls).pin_memory(); h_le = torch.cat(le).pin_memory()
h_out = torch.empty((n, 1), dtype=torch.float64).pin_memory(); h_ov = torch.empty((n, 1), dtype=torch.uint8).pin_memory()
off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
R = [(a.numpy().view(np.uint32), b.numpy().view(np.uint32), c.numpy()) for a, b, c in rights]
def run():
    ctx.map_whole(R, off, h_ls.numpy().view(np.uint32), h_le.numpy().view(np.uint32), ["mean"], out=h_out.numpy(), out_valid=h_ov.numpy())
for chunks, workers in [(8, 4), (22, 4), (22, 6), (12, 3), (22, 2), (44, 4), (22, 8)]:
    os.environ["GRB_WHOLE_CHUNKS"] = str(chunks); os.environ["GRB_WHOLE_WORKERS"] = str(workers)
    run(); best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"chunks={chunks:3d} workers={workers} {best*1e3:7.2f} ms", flush=True)
