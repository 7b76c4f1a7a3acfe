#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per captured launch the metrics the roofline discussion needs."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:60], r[idx['Block Size']], r[idx['Grid Size']])
    for w in want:
        if w in idx:
            print(f'  {w} = {r[idx[w]]} {units[idx[w]]}')
    items = []
    for h, i in idx.items():
        if 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h:
            try:
                items.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('  stalls/issue:', ', '.join(f'{h} {v:.2f}' for v, h in sorted(items, reverse=True)[:7]))
