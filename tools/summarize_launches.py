#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel launch count, total device time and share.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] > profiles/<name>.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.replace("void ", "").replace("vapb::", "").strip()[:60]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"],
                     float(r["Metric Value"].replace(",", "")) / 1e3))
    rows = [r for r in rows if r[0] >= skip]
    agg = OrderedDict()
    for _, k, grid, blk, us in rows:
        a = agg.setdefault(k, [0, 0.0, grid, blk])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}  launches: {len(rows)} (skipped first {skip})  total {tot / 1e3:.2f} ms "
          "(ncu-serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | total ms | share | avg us | grid (last) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, (n, us, grid, blk) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us / 1e3:.3f} | {us / tot * 100:.1f}% | {us / n:.1f} | {grid} | {blk} |")


if __name__ == "__main__":
    main()
