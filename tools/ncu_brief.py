#!/usr/bin/env python
"""Brief of one ncu --set full report: headline metrics + warp stall reasons per issue.  python tools/ncu_brief.py rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u, v = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum']
for k in keys:
    if k in h:
        print(f"{k:75s} {v[h.index(k)]:>16s} {u[h.index(k)]}")
st = []
for i, k in enumerate(h):
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
        try:
            st.append((float(v[i]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        except ValueError:
            pass
print("stall cycles per issued instruction:", ", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True) if x > 0.03))
