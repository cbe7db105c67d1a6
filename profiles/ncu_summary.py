#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion uses.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu",
        "sm__pipe_fp64_cycles_active", "smsp__warp_issue_stalled"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        print("==", name[:100])
        for i, n in enumerate(h):
            if any(n.startswith(k) for k in KEYS) and r[i] not in ("", "0"):
                if ".min" in n or ".max." in n or ".sum.p" in n and "bytes" not in n:
                    continue
                print(f"  {n:90s} {r[i]:>18s} {u[i]}")
    if "--source" in sys.argv:
        top = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hi = next(i for i, r in enumerate(rows) if "Source" in r or "# Source" in r or (r and r[0].startswith("#")))
        h = rows[hi]
        print(h)


if __name__ == "__main__":
    main()
