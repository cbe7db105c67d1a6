"""grb_map_joins_whole_f64_mgpu on all GPUs under different worker counts per GPU (fresh process per setting)."""
import os, sys, subprocess
code = r'''
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import granges_b200 as gb
n = 100_000_000
dev = torch.device("cuda", 0)
per = bench.split_counts(n, 22)
pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
rights, ls, le = [], [], []
for s in range(22):
    rs, re, sc = bench.gen_ranges(s, per[s], 14, dev, True)
    a, b = bench.gen_ranges(s, per[s], 13, dev, False)
    rights.append((pin(rs.cpu().numpy().view(np.uint32)), pin(re.cpu().numpy().view(np.uint32)), pin(sc.cpu().numpy())))
    ls.append(a.cpu().numpy().view(np.uint32)); le.append(b.cpu().numpy().view(np.uint32))
ls, le = pin(np.concatenate(ls)), pin(np.concatenate(le))
off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
out, ov = pin(np.empty((n, 1))), pin(np.empty((n, 1), np.uint8))
nd = torch.cuda.device_count()
ts = []
for rep in range(5):
    t0 = time.perf_counter()
    gb.map_whole_mgpu(list(range(nd)), rights, off, ls, le, ["mean"], out=out, out_valid=ov)
    ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("GRB_WHOLE_WORKERS"), os.environ.get("CUDA_DEVICE_SCHEDULE"), [round(t, 2) for t in ts[1:]], flush=True)
'''
for w in ("1", "2", "3", "4"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GRB_WHOLE_WORKERS=w))
