"""e2e (grb_map_joins_whole_f64, 100M x 100M from host buffers) under different worker / chunk counts; pinned and pageable."""
import os, sys, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for workers, chunks in ((4, 8), (6, 8), (8, 8), (4, 12), (6, 16), (8, 22)):
    env = dict(os.environ, GRB_WHOLE_WORKERS=str(workers), GRB_WHOLE_CHUNKS=str(chunks))
    r = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--no-configs"], capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(workers, chunks, "pinned", round(d["e2e"]["ms_per_step"], 2), "pageable", round(d["e2e"]["pageable"]["ms_per_step"], 2), flush=True)
    except Exception as e:
        print(workers, chunks, "failed", r.stderr[-300:])
