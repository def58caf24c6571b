#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        try:
            name = row[ki].split("(")[0]
            v = float(row[vi].replace(",", ""))
        except Exception:
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} n={n:5d} total={t / 1e6:9.3f} ms avg={t / n / 1e3:9.2f} us share={t / tot:6.1%}")
    print("total ms", tot / 1e6)


if __name__ == "__main__":
    main(sys.argv[1])
