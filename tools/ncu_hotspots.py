#!/usr/bin/env python
"""Warp-stall hot spots of one kernel from an `ncu --set full --import-source on` report (SASS view):
   python tools/ncu_hotspots.py report.ncu-rep [min_pct] > profiles/<name>_hotspots.txt"""
import csv
import subprocess
import sys


def main(rep, min_pct=1.0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # one block per kernel instance: "Kernel Name" row, header row, instruction rows
    k = 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name":
            name = rows[k][1]
            hdr = rows[k + 1]
            ci = {h: i for i, h in enumerate(hdr)}
            body = []
            k += 2
            while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
                if len(rows[k]) > ci["# Samples"]:
                    body.append(rows[k])
                k += 1
            tot = sum(float(r[ci["# Samples"]] or 0) for r in body) or 1.0
            inst = sum(float(r[ci["Instructions Executed"]] or 0) for r in body)
            print("# %s\n# samples %d, warp instructions %d; lines with >= %.1f %% of the samples" % (name[:100], tot, inst, min_pct))
            for n, r in enumerate(body):
                p = 100.0 * float(r[ci["# Samples"]] or 0) / tot
                if p >= min_pct:
                    print("%5d %6.2f%%  exec %10s  %s" % (n, p, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:100]))
        else:
            k += 1


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
