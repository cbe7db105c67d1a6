"""One grb_map_joins_whole_f64 call at full size with GRB_TIMING=1: the per-job wall-clock log (stderr) of the whole call."""
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
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize(); print("ms", (time.perf_counter() - t0) * 1e3, flush=True)
os.environ["GRB_TIMING"] = "1"
run()
# raw link rate of this process, for reference: 1 GiB page-locked -> device, three times
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); y = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize()
    print("raw H2D GB/s", round((1 << 30) / (time.perf_counter() - t0) / 1e9, 1), flush=True)
