#!/usr/bin/env python
"""Print the handful of ncu metrics this repo's roofline argument uses.
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_summary.py raw.csv"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'sm__inst_executed.sum.per_cycle_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__lsu_writeback_active_mem_lg.sum', 'sm__inst_executed_pipe_uniform.sum']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print('#', name[:100])
    for h, u, v in zip(hdr, units, vals):
        if h in keep or h.startswith('smsp__average_warps_issue_stalled') or h.startswith('smsp__inst_executed_pipe_'):
            if h.startswith('smsp__inst_executed_pipe_') and not h.endswith('.sum'):
                continue
            print(f"{h:92s} {u:16s} {v}")
