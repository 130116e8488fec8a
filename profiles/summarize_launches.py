#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, avg ms)."""
import collections
import csv
import re
import sys


def main(path, steps=1):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?(?:g[dh]::)?(\w+)<g[dh]::(\w+)", name)
        key = f"{m.group(1)}<{m.group(2)}>" if m else name.split("(")[0][:48]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v
        agg[key][0] += 1
        agg[key][1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"{'kernel':42s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:42s} {n:8d} {ms:10.2f} {ms / n:9.3f} {100 * ms / total:6.1f}%")
    print(f"{'sum (serialised, cold cache)':42s} {'':8s} {total:10.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
