#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout -k 10 400 python bench.py --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    grep "gp fine" gpurun_out/$name.err > gpurun_out/$name.fine; }
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=900 ) > gpurun_out/gpu_tests.log 2>&1 || echo "TESTS FAILED"
tail -8 gpurun_out/gpu_tests.log
run base_16M X=1 -- --workload uniform_16M --steps 20 --warmup 5
run base_2M X=1 -- --workload uniform_2M --steps 20 --warmup 5 --e2e-steps 0
run base_mixed X=1 -- --workload mixed_4M --steps 20 --warmup 5 --e2e-steps 0
run fine_2M GP_PROF_FINE=1 -- --workload uniform_2M --steps 20 --warmup 5 --e2e-steps 0
run fine_16M GP_PROF_FINE=1 -- --workload uniform_16M --steps 20 --warmup 5 --e2e-steps 0
python tools/show_bench.py gpurun_out/base_*.json | grep -v n_bodies
python - <<'P'
import json
d=json.loads(open('gpurun_out/base_16M.json').read().strip().splitlines()[-1])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "dependent", e["dependent"]["value"], [ (x["dirty_fraction"], x["value"], x["ms_per_step"]) for x in e["dirty_set"]])
P
grep "k_contacts" gpurun_out/fine_2M.fine | grep -v slot; grep "k_contacts /" gpurun_out/fine_16M.fine | grep -v slot
