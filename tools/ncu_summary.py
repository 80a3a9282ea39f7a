# -*- coding: utf-8 -*-
""" Condenses `ncu -i X.ncu-rep --page raw --csv` output into the handful of metrics quoted in DESIGN.md /
profiles/README.md.  Usage:  ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > summary.csv """
import csv
import sys

WANT = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_op_read_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
STALL = 'smsp__average_warps_issue_stalled_'

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')]
cols = [w for w in WANT if w in idx] + stalls
out = csv.writer(sys.stdout)
out.writerow([c.replace(STALL, 'stall_').replace('_per_issue_active.ratio', '') for c in cols])
out.writerow([units[idx[c]] for c in cols])
for r in data:
    out.writerow([r[idx[c]] for c in cols])
