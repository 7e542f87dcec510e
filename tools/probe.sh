#!/bin/bash
set -u
mkdir -p gpurun_out
run() { # name, env..., -- bench args
    name=$1; shift
    envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout -k 10 300 python bench.py --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    grep "gp fine" gpurun_out/$name.err > gpurun_out/$name.fine
}
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout=300 --deselect tests/test_gpu_large.py > gpurun_out/gpu_tests.log 2>&1 || echo "TESTS FAILED"
tail -5 gpurun_out/gpu_tests.log
run base_16M X=1 -- --workload uniform_16M --steps 20 --warmup 5
run base_2M X=1 -- --workload uniform_2M --steps 20 --warmup 5
run fine_16M GP_PROF_FINE=1 -- --workload uniform_16M --steps 20 --warmup 5 --e2e-steps 0
run fine_2M GP_PROF_FINE=1 -- --workload uniform_2M --steps 20 --warmup 5 --e2e-steps 0
for t in 512; do
  GP_NVCC_EXTRA="-DGP_COOP_THREADS=$t" python -c "from gears_b200 import build; build.build(force=True)" > gpurun_out/build_$t.log 2>&1
  run t${t}_16M X=1 -- --workload uniform_16M --steps 20 --warmup 5
  run t${t}_2M X=1 -- --workload uniform_2M --steps 20 --warmup 5
done
python tools/show_bench.py gpurun_out/base_*.json gpurun_out/t*_*.json | grep -v n_bodies
for f in fine_16M fine_2M; do grep "k_contacts\|per step\|slot" gpurun_out/$f.fine; done
