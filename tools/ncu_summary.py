#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into one block per kernel launch: duration, DRAM bytes, hit rates, top stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('=====', r[ki][:110])
    for i, h in enumerate(hdr):
        if h in want:
            print(f'  {h:75s} {units[i]:12s} {r[i]}')
    stalls = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active' in h:
            try: stalls.append((float(r[i]), h.split('issue_stalled_')[1].split('_per_issue')[0]))
            except ValueError: pass
    stalls.sort(reverse=True)
    print('  stalls:', ', '.join(f'{n}={v:.1f}' for v, n in stalls[:6]))
