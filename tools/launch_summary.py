#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
share of the total device time, launch count, total and mean duration.
Usage: tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(x for x in open(sys.argv[1]) if x.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(v[1] for v in agg.values())
    print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {total:.1f} ms of kernel time "
          "(ncu per-launch times: cold-cache, serialised — compare shares, not absolutes)")
    print(f"{'share':>7} {'launches':>9} {'total ms':>12} {'mean us':>11}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / total * 100:6.2f}% {v[0]:9d} {v[1]:12.3f} {v[1] / v[0] * 1e3:11.1f}  {k[:150]}")


if __name__ == "__main__":
    main()
