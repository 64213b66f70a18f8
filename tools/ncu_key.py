#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep, one block per kernel (quick look; tools/ncu_summary.py writes the profiles/ files)."""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.avg',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(w, r[hdr.index(w)], rows[1][hdr.index(w)])
    for h, v in zip(hdr, r):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3:
            print("  stall", h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v)
    print()
