#!/usr/bin/env python
"""Development aid: the handful of ncu raw-page metrics that matter for the issue-bound phase kernels.
usage: ncu_brief.py file.ncu-rep [...]"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_xu.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum', 'sm__inst_executed_pipe_uniform.sum']

for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        print('==', f, d.get('Kernel Name', ('?',))[0][:60])
        for h in hdr:
            if h in KEYS or ('smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h):
                v, u = d[h]
                try:
                    if 'stalled' in h and float(v) < 0.1:
                        continue
                except ValueError:
                    pass
                print(f'  {h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""):70s} {v} {u}')
