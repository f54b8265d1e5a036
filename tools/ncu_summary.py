#!/usr/bin/env python3
"""Extracts the metrics DESIGN.md / profiles/ quote from an .ncu-rep (run where ncu is installed):
   python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_ncu_x.csv"""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'launch__occupancy_limit_shared_mem', 'sm__maximum_warps_per_active_cycle_pct',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__sass_thread_inst_executed_op_integer_pred_on.sum', 'sm__sass_thread_inst_executed_op_integer_pred_on.sum']
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
for r in rows[2:]:
    w.writerow(['metric', 'unit', 'value'])
    for i, h in enumerate(hdr):
        if h in KEEP or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')) \
                or (h.startswith('smsp__average_warp_latency_issue_stalled') and h.endswith('.ratio')):
            w.writerow([h, units[i], r[i]])
