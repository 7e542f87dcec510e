#!/bin/bash
# quick single-GPU check: all GPU tests, then the kernel-only line at 16 M and 2 M bodies
set -u
mkdir -p gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=900 ) > gpurun_out/gpu_tests.log 2>&1 || echo "TESTS FAILED"
tail -6 gpurun_out/gpu_tests.log
for wl in uniform_16M uniform_2M mixed_4M; do
python bench.py --workload $wl --no-cpu --e2e-steps 0 > gpurun_out/q_$wl.json 2> gpurun_out/q_$wl.err
python tools/show_bench.py gpurun_out/q_$wl.json | grep -v n_bodies | cut -c1-500
done
