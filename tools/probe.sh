#!/bin/bash
# scratch probe for one gpurun call: GPU test suite + timing variants (results under gpurun_out/)
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/gpu_tests.log 2>&1
rc=$?
tail -15 gpurun_out/gpu_tests.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED rc=$rc"; fi
run() { # name, env..., -- bench args
    name=$1; shift
    envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout -k 10 300 python bench.py --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    grep "gp fine" gpurun_out/$name.err > gpurun_out/$name.fine
}
run base_16M X=1 -- --workload uniform_16M --steps 20 --warmup 5
run base_2M X=1 -- --workload uniform_2M --steps 20 --warmup 5
run fine_16M GP_PROF_FINE=1 -- --workload uniform_16M --steps 20 --warmup 5
run fine_2M GP_PROF_FINE=1 -- --workload uniform_2M --steps 20 --warmup 5
run fine_pile GP_PROF_FINE=1 -- --workload pile_1M --steps 10 --warmup 300
run base_mixed X=1 -- --workload mixed_4M --steps 20 --warmup 5
python tools/show_bench.py gpurun_out/base_*.json gpurun_out/fine_*.json
cat gpurun_out/fine_16M.fine gpurun_out/fine_2M.fine gpurun_out/fine_pile.fine
tail -3 gpurun_out/*.err | head -60
