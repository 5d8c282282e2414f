#!/usr/bin/env python3
"""Markdown table of the key metrics of `ncu --set full` captures. usage: ncu_summary.py rep [rep...]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
print("|---|" + "---|" * len(WANT))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("n/a")
        print(f"| {name} | " + " | ".join(cells) + " |")
