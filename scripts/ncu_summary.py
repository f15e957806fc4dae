#!/usr/bin/env python
"""Markdown table of the metrics profiles/ncu_kfields_r1.md quotes, from `ncu -i X.ncu-rep --page raw --csv`.
usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py raw.csv"""
import csv
import sys

WANT = """dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum launch__grid_size launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__registers_per_thread sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active""".split()
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
print("| metric | value | unit |\n|---|---|---|")
for i, h in enumerate(hdr):
    if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) >= 0.03):
        print("| %s | %s | %s |" % (h, vals[i], units[i]))
