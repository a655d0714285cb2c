#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list: per-kernel launch count, total device time, share and (when captured) DRAM traffic.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] [--each] > profiles/<name>.md
    python tools/summarize_launches.py gpurun_out/launches.csv --skip N --traffic-json profiles/<name>.json --steps K
(--traffic-json: DRAM bytes and ncu time per kernel family and per bench step, the file bench.py reads for
`roofline.traffic`; families as in csrc/model.h ProfCat, the downsample GEMM — the first gemm_lin after the
recurrence — counted with the conv GEMMs as the live profile does.)
"""
import csv
import re
import sys
from collections import OrderedDict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "").replace("vapb::", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.strip()[:60]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    launches = OrderedDict()
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        d = launches.setdefault(i, {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = v
    rows = [d for i, d in launches.items() if i >= skip]
    have_dram = any("dram__bytes_read.sum" in d for d in rows)
    if "--traffic-json" in sys.argv:
        import json

        out = sys.argv[sys.argv.index("--traffic-json") + 1]
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
        fam_of = {"conv0_tc_kernel": "conv0", "conv01_kernel": "conv_gemm", "gemm_2sm_kernel": "conv_gemm", "rnn_tc_kernel": "rnn",
                  "attention_tc_kernel": "attention", "gemm_lin_kernel": "linear_gemm", "ffn_fused_kernel": "linear_gemm",
                  "probs_kernel": "heads", "head_probs_kernel": "heads", "loss_kernel": "heads", "vad_head_blocked_kernel": "heads"}
        fams, after_rnn = {}, False
        for d in rows:
            fam = fam_of.get(d["name"], "other")
            if d["name"] == "rnn_tc_kernel":
                after_rnn = True
            elif after_rnn and d["name"] == "gemm_lin_kernel":
                fam, after_rnn = "conv_gemm", False
            f = fams.setdefault(fam, {"dram_bytes_per_step": 0.0, "launches": 0, "ncu_ms": 0.0})
            f["dram_bytes_per_step"] += (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / steps
            f["launches"] += 1
            f["ncu_ms"] += d.get("gpu__time_duration.sum", 0.0) / 1e3 / steps
        for f in fams.values():
            f["launches"] = f["launches"] // steps
            f["ncu_ms"] = round(f["ncu_ms"], 3)
        fams["total"] = {"dram_bytes_per_step": sum(f["dram_bytes_per_step"] for f in fams.values()),
                         "launches": sum(f["launches"] for f in fams.values()),
                         "ncu_ms": round(sum(f["ncu_ms"] for f in fams.values()), 3)}
        tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "B=256, 1 GPU"
        fams["_source"] = (f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over {steps} "
                           f"step(s) of bench.py ({tag}; {path}, first {skip} launches skipped)")
        json.dump(fams, open(out, "w"), indent=1)
        print(json.dumps(fams, indent=1))
        return
    if "--each" in sys.argv:
        print("| # | kernel | grid | us | dram rd MB | dram wr MB | GB/s |")
        print("|---|---|---|---:|---:|---:|---:|")
        for i, d in enumerate(rows):
            us = d.get("gpu__time_duration.sum", 0.0)
            rd, wr = d.get("dram__bytes_read.sum", 0.0) / 1e6, d.get("dram__bytes_write.sum", 0.0) / 1e6
            print(f"| {i} | {d['name']} | {d['grid']} | {us:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / us * 1e3 if us else 0:.0f} |")
        return
    agg = OrderedDict()
    for d in rows:
        a = agg.setdefault(d["name"], [0, 0.0, d["grid"], d["block"], 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[4] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}  launches: {len(rows)} (skipped first {skip})  total {tot / 1e3:.2f} ms "
          "(ncu-serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | total ms | share | avg us | grid (last) | block |" + (" DRAM GB | GB/s |" if have_dram else ""))
    print("|---|---:|---:|---:|---:|---|---|" + ("---:|---:|" if have_dram else ""))
    for k, (n, us, grid, blk, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = f" {by / 1e9:.2f} | {by / us / 1e3:.0f} |" if have_dram else ""
        print(f"| {k} | {n} | {us / 1e3:.3f} | {us / tot * 100:.1f}% | {us / n:.1f} | {grid} | {blk} |" + extra)


if __name__ == "__main__":
    main()
