#!/usr/bin/env python3
"""Summarises an .ncu-rep (ncu --set full) into a small markdown table.
usage: summarize_ncu.py report.ncu-rep > summary.md"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "L2 rd miss sectors"),
    ("lts__t_sectors_srcunit_tex_op_atom.sum", "L2 atom sectors"),
    ("dram__sectors_read.sum", "dram rd sectors"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall/issue"),
]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS if m in idx]
    print("| kernel | grid | " + " | ".join(n for _, n in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for m, _ in cols:
            v = r[idx[m]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[idx[m]]}".strip())
        print(f"| {name} | {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
