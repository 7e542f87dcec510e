#!/bin/bash
# Evaluate the opt-in experimental kernels on a GPU box, in ONE gpurun call:
#   gpurun --timeout 700 -- 'bash tools/try_experiments.sh'
# 1. parity of every variant against the oracle (no -x: every variant gets its verdict);
# 2. device-resident throughput of every variant on uniform_16M and uniform_2M next to the default kernels.
set -u
mkdir -p gpurun_out
GP_EXPERIMENTAL=1 timeout 330 python -m pytest tests/test_gpu_parity.py -q --timeout=100 -k experimental \
    > gpurun_out/exp_tests.log 2>&1
tail -25 gpurun_out/exp_tests.log
for v in base "GP_PAIRFIND=warp" "GP_PAIRFIND=warp256" "GP_WIDE_SCATTER=1" "GP_SWEEP_SKIP=1" \
         "GP_PAIRFIND=warp256 GP_WIDE_SCATTER=1 GP_SWEEP_SKIP=1"; do
    name=$(echo "$v" | tr ' =' '__')
    for wl in uniform_16M uniform_2M; do
        if [ "$v" = base ]; then
            timeout 60 python bench.py --workload $wl --no-cpu --steps 20 --warmup 5 \
                > gpurun_out/exp_${name}_${wl}.json 2> gpurun_out/exp_${name}_${wl}.err
        else
            env $v timeout 60 python bench.py --workload $wl --no-cpu --steps 20 --warmup 5 \
                > gpurun_out/exp_${name}_${wl}.json 2> gpurun_out/exp_${name}_${wl}.err
        fi
    done
done
python tools/show_bench.py gpurun_out/exp_*.json | grep -v n_bodies
