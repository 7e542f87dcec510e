#!/usr/bin/env python
"""Key metrics per captured launch from an .ncu-rep: python tools/ncu_raw.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__registers_per_thread', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
print(" | ".join(f"{w.split('.')[0][-28:]}[{units[i]}]" for w, i in idx))
for r in rows[2:]:
    print(" | ".join(r[i][:26] for _, i in idx))
if len(sys.argv) > 2:  # list metric names matching a substring
    for h in hdr:
        if sys.argv[2] in h: print(h)
