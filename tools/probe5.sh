#!/bin/bash
# one `ncu --set full` capture of k_find_pairs (the report comes back in gpurun_out/)
set -u
mkdir -p gpurun_out
CMD="python bench.py --no-cpu --steps 3 --warmup 3 --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_find_pairs -s 4 -c 1 -f -o gpurun_out/ncu_find_pairs $CMD > gpurun_out/ncu_fp.log 2>&1
tail -3 gpurun_out/ncu_fp.log; ls -la gpurun_out/*.ncu-rep
