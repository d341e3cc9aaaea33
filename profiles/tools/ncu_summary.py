#!/usr/bin/env python
"""Key metrics and the top stall reasons of every kernel in an `ncu --page raw --csv` export.
    ncu -i X.ncu-rep --page raw --csv > X.csv ; python profiles/tools/ncu_summary.py X.csv"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.per_cycle_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.max']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:100])
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f'  {h:72s} {r[i]} {units[i]}')
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and r[i]]
    st.sort(reverse=True)
    print('  stalls (warps per issue): ' + ', '.join(
        f'{h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]} {v:.2f}' for v, h in st[:7]))
