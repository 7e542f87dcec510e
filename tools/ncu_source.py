#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel: python tools/ncu_source.py rep.ncu-rep <kernel-regex> [topN]"""
import csv, subprocess, sys, io, collections
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kname,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
hdr = rows[hi]
ln, src = hdr.index('Line No'), hdr.index('Source')
sa, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
agg = collections.OrderedDict(); ts = ti = 0.0
cur = ('?', '')
for r in rows[hi + 1:]:
    if len(r) <= max(sa, ie): continue
    if r[ln].strip(): cur = (r[ln], r[src])
    try: s = float(r[sa] or 0); i = float(r[ie] or 0)
    except ValueError: continue
    a = agg.setdefault(cur, [0.0, 0.0]); a[0] += i; a[1] += s; ts += s; ti += i
print(f"total warp-instructions {ti:.0f}, stall samples {ts:.0f}")
for (l, sline), (i, s) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{100*i/max(ti,1):5.1f}% inst {100*s/max(ts,1):5.1f}% stall  L{l}: {sline.strip()[:110]}")
