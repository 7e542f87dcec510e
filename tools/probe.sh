#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=900 ) > gpurun_out/gpu_tests.log 2>&1 || echo "TESTS FAILED"
tail -12 gpurun_out/gpu_tests.log
CMD="python bench.py --workload uniform_16M --no-cpu --steps 4 --warmup 4 --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_find_pairs|k_scatter_cells|k_keys|k_sort_chains|k_fill_chains" -s 25 -c 5 -o gpurun_out/prof_main $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
ls -la gpurun_out/prof_main.ncu-rep
