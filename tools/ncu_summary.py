#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per profiled launch with the
counters the roofline discussion needs."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("lts__t_bytes.sum", "l2_MB", 1e-6),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("smsp__inst_executed.sum", "inst", 1),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit", 1),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {n: i for i, n in enumerate(hdr)}
    print("| # | kernel | grid | " + " | ".join(k[1] for k in KEYS if k[0] in col) + " |")
    print("|---|---|---|" + "---|" * sum(1 for k in KEYS if k[0] in col))
    for r in rows[hdr_i + 2:]:
        if not r or not r[0].isdigit():
            continue
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1][:28]
        grid = r[col["Grid Size"]].replace(" ", "") if "Grid Size" in col else ""
        vals = []
        for k, _, sc in KEYS:
            if k not in col:
                continue
            v = r[col[k]].replace(",", "")
            try:
                x = float(v)
                u = units[col[k]]
                if k == "gpu__time_duration.sum":
                    x = x * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
                    vals.append(f"{x:.1f}")
                elif k.startswith("dram__bytes") or k.startswith("lts__t_bytes"):
                    x = x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
                    vals.append(f"{x:.1f}")
                else:
                    vals.append(f"{x:.4g}")
            except ValueError:
                vals.append(v)
        print(f"| {r[0]} | {name} | {grid} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
