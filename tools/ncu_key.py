#!/usr/bin/env python
"""(no GPU) the handful of counters a kernel decision rests on, from one .ncu-rep: python tools/ncu_key.py <rep> [more metrics...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = dict(zip(rows[0], zip(rows[1], rows[2])))
for k in KEYS + sys.argv[2:]:
    if k in d:
        print(f"{k:90s} {d[k][1]:>16s} {d[k][0]}")
for k in sorted(d):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        try:
            if float(d[k][1]) >= 0.3:
                print(f"{k:90s} {d[k][1]:>16s}")
        except ValueError:
            pass
