#!/bin/bash
set -u
mkdir -p gpurun_out
for lvl in 1 3 1 3; do
GP_PROF_FINE=$lvl timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2974$lvl bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 0 --parity-steps 0 > gpurun_out/bench_2gpu_fine$lvl.json 2> gpurun_out/bench_2gpu_fine$lvl.err
echo "level $lvl"; python tools/show_bench.py gpurun_out/bench_2gpu_fine$lvl.json | head -2 | cut -c1-400
grep "slab proof" gpurun_out/bench_2gpu_fine$lvl.err
done
