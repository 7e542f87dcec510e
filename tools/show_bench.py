#!/usr/bin/env python
"""Compact view of bench.py JSON lines: python tools/show_bench.py gpurun_out/bench_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    print(f"{f}: value {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.3f} ms/step  gpus={d['n_gpus']} exact={d.get('exact')} launches={d.get('gpu_launches')}")
    if d.get('kernels'):
        print("   ", {k: (v['ms'], v['gbs']) for k, v in d['kernels'].items()})
    if d.get('step_stats'): print("   ", d['step_stats'])
    if d.get('e2e'): print("    e2e %.4g" % d['e2e']['value'], " cpu", d.get('cpu_baseline') and d['cpu_baseline'].get('value'), " roofline", d.get('roofline') and (d['roofline'].get('kernel'), d['roofline']['frac']))
