"""Phases of the host-buffer (e2e) path at full size: raw pinned copies, grb_index_build_batch, grb_map_joins_f64_batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, granges_b200 as gb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
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
d = torch.empty(n * 2, dtype=torch.int32, device=dev)
def T(name, f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); ctx.sync(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{name:58s} {best*1e3:8.2f} ms", flush=True); return r
T("raw H2D 0.8 GB pinned (lefts)", lambda: (d[:n].copy_(h_ls, non_blocking=True), d[n:].copy_(h_le, non_blocking=True)))
dd = torch.empty(n, dtype=torch.float64, device=dev)
T("raw D2H 0.8 GB pinned (results)", lambda: h_out.view(-1).copy_(dd, non_blocking=True))
hs = None
def build():
    global hs
    if hs:
        for h in hs: h.close()
    hs = ctx.index_build_batch([(a.numpy().view(np.uint32), b.numpy().view(np.uint32), c.numpy()) for a, b, c in rights])
T("grb_index_build_batch (H2D 1.6 GB + 22 builds)", build)
T("grb_map_joins_f64_batch host in/out (H2D 0.8, map, D2H 0.9)", lambda: ctx.map_joins_batch(hs, off, h_ls.numpy().view(np.uint32), h_le.numpy().view(np.uint32), ["mean"], out=h_out.numpy(), out_valid=h_ov.numpy()))
dl, de = h_ls.to(dev), h_le.to(dev)
do = torch.empty((n, 1), dtype=torch.float64, device=dev); dv = torch.empty((n, 1), dtype=torch.uint8, device=dev)
T("map_joins_batch device in/out", lambda: ctx.map_joins_batch(hs, off, dl, de, ["mean"], out=do, out_valid=dv))
drs = [tuple(t.to(dev) for t in r) for r in rights]
def build_dev():
    xs = [ctx.index_build(a, b).set_scores(c) for a, b, c in drs]
    for x in xs: x.close()
T("22 index builds, device-resident inputs, one stream", build_dev)
