#!/usr/bin/env python
"""Top source lines by warp-stall samples: python tools/ncu_source.py rep.ncu-rep <kernel-substr> [launch-index]"""
import csv, subprocess, sys, io, collections, re
rep, kname = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kname] +
                     (["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else ["--launch-count", "1"]),
                     capture_output=True, text=True).stdout
lines = out.splitlines()
# find header row
start = next(i for i, l in enumerate(lines) if l.startswith('"') and 'Source' in l)
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
print(hdr[:12])
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name: return i
    return None
ci_src = col('Source'); ci_samp = col('# Samples') or col('Warp Stall Sampling (All Samples)') or col('Samples')
print('cols', ci_src, ci_samp)
agg = collections.Counter(); cur = None
tot = 0
for r in rows[1:]:
    if len(r) <= max(ci_src, ci_samp or 0): continue
    try: v = float(r[ci_samp].replace(',', '') or 0)
    except: v = 0
    agg[r[ci_src][:110]] += v; tot += v
for k, v in agg.most_common(40):
    print(f"{v:8.0f} {100*v/max(1,tot):5.1f}%  {k}")
