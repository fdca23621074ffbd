#!/usr/bin/env python3
"""Aggregate an ncu launch list (``--metrics gpu__time_duration.sum --csv``) by kernel: launches, total, mean, share."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        us = float(row["Metric Value"].replace(",", "")) * scale[row["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"{'launches':>8} {'total ms':>10} {'mean us':>10} {'share':>7}  kernel")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:8d} {us / 1e3:10.3f} {us / n:10.1f} {us / total:7.1%}  {name}")
    print(f"{sum(a[0] for a in agg.values()):8d} {total / 1e3:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
