#!/usr/bin/env python
"""Aggregate an `ncu --csv --metrics ...` launch list per kernel: count, mean of every metric.

  python tools/ncu_agg.py gpurun_out/launches.csv [--by-grid]
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def main():
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    h = rows[0]
    ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[iid], r[ik].split("(")[0]), {})[r[im]] = float(r[iv].replace(",", ""))
    agg = defaultdict(lambda: defaultdict(list))
    for (_, name), met in per.items():
        key = name if not by_grid else "%s grid=%d" % (name, int(met.get("launch__grid_size", 0)))
        for a, b in met.items():
            agg[key][a].append(b)
    short = lambda a: (a.replace("smsp__average_warps_issue_stalled_", "st_").replace("_per_issue_active.ratio", "")
                       .replace("smsp__", "").replace(".avg.pct_of_peak_sustained_active", "%")
                       .replace("gpu__time_duration.sum", "ns").replace("launch__grid_size", "grid"))
    tot = sum(sum(v.get("gpu__time_duration.sum", [0])) for v in agg.values())
    for key, met in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
        n = len(met.get("gpu__time_duration.sum", [0]))
        t = sum(met.get("gpu__time_duration.sum", [0]))
        print("%-34s n=%4d total=%9.1f us share=%5.1f%% | %s" % (
            key[:34], n, t / 1e3, 100.0 * t / max(tot, 1),
            " ".join("%s=%.4g" % (short(a), sum(b) / len(b)) for a, b in met.items())))


if __name__ == "__main__":
    main()
