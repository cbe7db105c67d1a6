"""Where the time of grb_map_joins_whole_f64_mgpu goes (GRB_TIMING=1 for the per-job log)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for devices in [[0], list(range(nd))]:
    for rep in range(3):
        if rep == 2: os.environ["GRB_TIMING"] = "1"
        t0 = time.perf_counter()
        gb.map_whole_mgpu(devices, rights, off, ls, le, ["mean"], out=out, out_valid=ov)
        print(devices, rep, round((time.perf_counter() - t0) * 1e3, 2), "ms", flush=True)
        os.environ.pop("GRB_TIMING", None)
