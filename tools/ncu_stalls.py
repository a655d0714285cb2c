#!/usr/bin/env python
"""Per-kernel stall summary from `ncu -i X.ncu-rep --page source --csv`: total samples by stall reason
and the top-N instructions by samples."""
import csv
import sys


def main(path, topn=25):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(starts, starts[1:]):
        print("==", rows[a][1][:100])
        hdr = rows[a + 1]
        col = {n: i for i, n in enumerate(hdr)}
        data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
        sc = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        tot = {n: sum(int(r[col[n]]) for r in data if r[col[n]].isdigit()) for n in sc}
        all_s = sum(tot.values())
        print("instructions:", len(data), " executed:", sum(int(r[col["Instructions Executed"]]) for r in data),
              " samples:", all_s)
        print("  " + "  ".join(f"{n[6:]}={v} ({100 * v / all_s:.0f}%)" for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
        for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:topn]:
            st = sorted(((int(r[col[n]]), n[6:]) for n in sc if r[col[n]].isdigit()), reverse=True)[:2]
            print(f"  {r[col['# Samples']]:>6} {r[col['Source']].strip()[:80]:<80} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
