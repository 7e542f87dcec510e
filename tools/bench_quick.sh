#!/bin/bash
# usage: tools/bench_quick.sh [bench args]  -> compact summary of one bench run
python bench.py "$@" --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g body-steps/s  %.3f ms/step  exact=%s launches=%d' % (d['value'], d['ms_per_step'], d['exact'], d['gpu_launches']))
print({k:v['ms'] for k,v in d['kernels'].items()})
print(d['step_stats'])
print('e2e', d['e2e'] and d['e2e']['value'])"
