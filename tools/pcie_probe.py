"""Raw concurrent host-to-device bandwidth from page-locked memory to every visible GPU (are the links independent?)."""
import time, torch, subprocess, threading
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
n = torch.cuda.device_count()
nb = 1 << 30
hosts = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(n)]
devs = [torch.empty(nb, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
def run(ids, d2h=False):
    streams = [torch.cuda.Stream(device=i) for i in ids]
    for i in ids: torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for rep in range(4):
        for i, s in zip(ids, streams):
            with torch.cuda.stream(s):
                if d2h: hosts[i].copy_(devs[i], non_blocking=True)
                else: devs[i].copy_(hosts[i], non_blocking=True)
    for i in ids: torch.cuda.synchronize(i)
    dt = time.perf_counter() - t0
    return 4 * nb * len(ids) / dt / 1e9
for ids in [[0]] + ([[1], [0, 1]] if n > 1 else []) + ([list(range(n))] if n > 2 else []):
    print("H2D", ids, round(run(ids), 1), "GB/s   D2H", round(run(ids, True), 1), "GB/s", flush=True)
