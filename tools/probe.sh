#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=900 --durations=5 ) > gpurun_out/gpu_tests.log 2>&1 || echo "TESTS FAILED"
tail -14 gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
python tools/show_bench.py gpurun_out/bench_default.json | grep -v n_bodies
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "dependent", e["dependent"]["value"])
for x in e["dirty_set"]: print(x["dirty_fraction"], x["value"], x["ms_per_step"], sum(x["step_groups_ms"].values()))
print(d["roofline"]); print(d["cpu_baseline"]["value"], d["cpu_baseline_grid"]["value"])
P
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -3 gpurun_out/bench_reference.err; cut -c1-600 gpurun_out/bench_reference.json
CMD="python bench.py --no-cpu --steps 3 --warmup 3 --e2e-steps 2"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_uniform_16M.csv $CMD > gpurun_out/ncu.log 2>&1
python tools/ncu_launches.py gpurun_out/ncu_launches_uniform_16M.csv | head -14
