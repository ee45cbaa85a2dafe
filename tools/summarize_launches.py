#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total, share, average.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [--last-step N_LAUNCHES]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    last = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[2] == "--last-step" else None
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name).replace("<unnamed>::", "")
        rows.append((name, float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"], r["Block Size"]))
    if last:
        rows = rows[-last:]
    agg = OrderedDict()
    for name, us, grid, blk in rows:
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += us
    tot = sum(d[1] for d in agg.values())
    print(f"total {tot / 1e3:.2f} ms over {len(rows)} launches (cold-cache, serialised under ncu: compare shares)")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:62s} {n:5d} {us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  avg {us / n:9.1f} us")


if __name__ == "__main__":
    main()
