"""BASELINE.json config C5 on N GPUs: full join-pair materialisation, 50M GTF-like lefts x 200M reads-like rights,
seqnames dealt to the ranks (longest-processing-time first on the expected pairs), no collective on the data path.
usage: torchrun --nproc-per-node N tools/bench_c5_mgpu.py [--scale F]      (N = 1: plain python)"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import granges_b200 as gb
from granges_b200 import _capi

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist

    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    os.dup2(saved, 1)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = gb.Context(local, stream=st)
HG38 = bench.HG38
nl, nr = int(50_000_000 * scale), int(200_000_000 * scale)
nl_per, nr_per = bench.split_counts(nl, 22), bench.split_counts(nr, 22)
# expected pairs of a seqname ~ NL * NR * (mean left length + mean right length) / seqname length
cost = [nl_per[s] * nr_per[s] / HG38[s] for s in range(22)]
load = [0.0] * world
mine = []
for s in sorted(range(22), key=lambda s: -cost[s]):
    r = int(np.argmin(load))
    load[r] += cost[s]
    if r == rank:
        mine.append(s)
L = _capi.lib()
data = []
for s in sorted(mine):
    g = torch.Generator(device=dev)
    g.manual_seed(500 + s)
    ln = torch.exp(np.log(2000.0) + torch.randn(nl_per[s], generator=g, device=dev, dtype=torch.float64)).clamp(50, 2e6).to(torch.int64)
    stt = (torch.rand(nl_per[s], generator=g, device=dev, dtype=torch.float64) * (HG38[s] - ln).to(torch.float64)).to(torch.int64)
    key, _ = torch.sort(stt * 4294967296 + (stt + ln))
    ls, le = (key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32)
    rl = torch.randint(75, 152, (nr_per[s],), generator=g, device=dev)
    rst = (torch.rand(nr_per[s], generator=g, device=dev, dtype=torch.float64) * (HG38[s] - 151)).to(torch.int64)
    key, _ = torch.sort(rst * 4294967296 + (rst + rl))
    ix = ctx.index_build((key >> 32).to(torch.int32), (key & 0xFFFFFFFF).to(torch.int32))
    data.append((ix, ls, le, torch.empty(nl_per[s] + 1, dtype=torch.int64, device=dev)))
ctx.sync()


def one_pass():
    pairs = 0
    for ix, ls, le, off in data:
        ph = C.c_void_p()
        rc = L.grb_left_overlaps(ctx.h, ix.h, ls.data_ptr(), le.data_ptr(), len(ls), off.data_ptr(), C.byref(ph))
        assert rc == 0, ctx._check(rc)
        pairs += L.grb_pairs_len(ph)
        L.grb_pairs_destroy(ph)
    ctx.sync()
    return pairs


one_pass()
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
pairs = one_pass()
torch.cuda.synchronize()
t = torch.tensor([time.perf_counter() - t0, float(pairs)], dtype=torch.float64, device=dev)
tmax = t.clone()
if dist is not None:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
if rank == 0:
    ms = float(tmax[0].item()) * 1e3
    P = int(t[1].item())
    print(json.dumps({"config": "C5 left_overlaps pairs %d x %d" % (nl, nr), "n_gpus": world, "ms": ms, "pairs": P, "pairs_per_s": P / ms * 1e3,
                      "lefts_per_s": nl / ms * 1e3, "note": "wall clock of the second pass, max over ranks; per seqname: count, scan, emit, "
                      "allocation from the pool and the pair-count read-back"}), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
