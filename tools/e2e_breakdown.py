"""Where does the host-buffer (e2e) path spend its time?  One seqname-sized index, pinned host buffers."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, granges_b200 as gb

use_default = len(sys.argv) > 1 and sys.argv[1] == "default"
use_torch_stream = len(sys.argv) > 1 and sys.argv[1] == "torchstream"
n = 4_545_455
rs, re, sc = bench.gen_ranges(0, n, 14, "cuda", True)
ls, le = bench.gen_ranges(0, n, 13, "cuda", False)
h = [t.cpu().pin_memory() for t in (rs, re, sc, ls, le)]
hrs, hre, hls, hle = (h[i].numpy().view(np.uint32) for i in (0, 1, 3, 4))
hsc = h[2].numpy()
out = torch.empty((n, 1), dtype=torch.float64).pin_memory().numpy()
ov = torch.empty((n, 1), dtype=torch.uint8).pin_memory().numpy()
if use_torch_stream:
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = gb.Context(0, stream=torch.cuda.current_stream() if (use_default or use_torch_stream) else None)
def T(f, name, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); ctx.sync(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{name:40s} {best*1e3:9.3f} ms"); return r
ix = T(lambda: ctx.index_build(hrs, hre), "index_build (host pinned)")
T(lambda: ix.set_scores(hsc), "set_scores (host pinned)")
T(lambda: ctx.map_joins(ix, hls, hle, ["mean"]), "map_joins alloc outputs (host)")
off = np.array([0, n], np.uint64)
T(lambda: ctx.map_joins_batch([ix], off, hls, hle, ["mean"], out=out, out_valid=ov), "map_joins_batch pinned in/out")
ix2 = T(lambda: ctx.index_build(rs, re), "index_build (device)")
T(lambda: ix2.set_scores(sc), "set_scores (device)")
T(lambda: ix2.close(), "index close", reps=1)
T(lambda: torch.empty(n * 2, dtype=torch.int32, device="cuda").copy_(torch.from_numpy(np.concatenate([hrs, hre]).view(np.int32)).pin_memory(), non_blocking=True), "torch H2D 36MB pinned")
