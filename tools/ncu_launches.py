#!/usr/bin/env python
"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list per kernel."""
import csv, collections, sys
def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row['Kernel Name'].split('(')[0].replace('void ', '')
        m = row['Metric Name']; v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
        a = agg.setdefault(k, dict(n=0, us=0.0, rd=0.0, wr=0.0))
        if m == 'gpu__time_duration.sum':
            v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}.get(u, v)  # to us
            a['n'] += 1; a['us'] += v
        elif m.startswith('dram__bytes'):
            v = {'byte': v, 'Kbyte': v * 1e3, 'Mbyte': v * 1e6, 'Gbyte': v * 1e9}.get(u, v)
            a['rd' if 'read' in m else 'wr'] += v
    return agg
def to_json(agg, workload, source):
    """profiles/ncu_traffic.json entry: per kernel, DRAM bytes (read + write) and microseconds per launch."""
    import json
    return {"source": source, "kernels": {k: {"launches": a['n'], "us_per_launch": round(a['us'] / max(1, a['n']), 1),
                                               "dram_read_bytes_per_launch": round(a['rd'] / max(1, a['n'])),
                                               "dram_write_bytes_per_launch": round(a['wr'] / max(1, a['n'])),
                                               "dram_bytes_per_launch": round((a['rd'] + a['wr']) / max(1, a['n']))}
                                           for k, a in agg.items()}}
if __name__ == '__main__':
    if '--json' in sys.argv:  # python tools/ncu_launches.py <csv> --json <workload> <out.json>: merge into the traffic file
        import json, os
        csvp, wl, outp = sys.argv[1], sys.argv[3], sys.argv[4]
        d = json.load(open(outp)) if os.path.exists(outp) else {"workloads": {}}
        d["workloads"][wl] = to_json(load(csvp), wl, csvp)
        json.dump(d, open(outp, 'w'), indent=1)
        sys.exit(0)
    agg = load(sys.argv[1]); steps = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != '--json' else None
    tot = sum(a['us'] for a in agg.values())
    print(f"{sum(a['n'] for a in agg.values())} launches, {tot:.1f} us total" + (f", {tot/steps:.1f} us/step over {steps:g} steps" if steps else ""))
    for k, a in sorted(agg.items(), key=lambda x: -x[1]['us']):
        print(f"{k[:44]:44s} n={a['n']:4d} total {a['us']:10.1f} us avg {a['us']/max(1,a['n']):9.1f} us share {a['us']/tot:.3f}"
              f"  dram R {a['rd']/max(1,a['n'])/1e6:8.1f} MB W {a['wr']/max(1,a['n'])/1e6:8.1f} MB per launch")
