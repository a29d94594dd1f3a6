#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg.setdefault(row["Kernel Name"][:72], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'avg us':>10s} {'total ms':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} {len(v):5d} {sum(v) / len(v):10.1f} {sum(v) / 1e3:9.2f} {100 * sum(v) / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
