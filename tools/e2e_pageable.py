"""grb_map_joins_whole_f64 at full size on ordinary (pageable) numpy arrays: the library's page-locked staging rings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, granges_b200 as gb

n = 100_000_000
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = gb.Context(0, stream=st)
per = bench.split_counts(n, 22)
R, ls, le = [], [], []
for s in range(22):
    a, b = bench.gen_ranges(s, per[s], 13, dev, False)
    rs, re, sc = bench.gen_ranges(s, per[s], 14, dev, True)
    R.append((rs.cpu().numpy().view(np.uint32), re.cpu().numpy().view(np.uint32), sc.cpu().numpy()))
    ls.append(a.cpu().numpy()); le.append(b.cpu().numpy())
p_ls = np.concatenate(ls).view(np.uint32); p_le = np.concatenate(le).view(np.uint32)
p_out = np.empty((n, 1), np.float64); p_ov = np.empty((n, 1), np.uint8)
off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
for _ in range(4):
    t0 = time.perf_counter(); ctx.map_whole(R, off, p_ls, p_le, ["mean"], out=p_out, out_valid=p_ov); print("pageable ms", round((time.perf_counter() - t0) * 1e3, 2), flush=True)
