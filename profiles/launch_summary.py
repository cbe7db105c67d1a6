#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for n, a in agg.items():
        print(f"{n:44s} {a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e3:10.1f} {a[1] / tot * 100:6.1f}%")


if __name__ == "__main__":
    main()
