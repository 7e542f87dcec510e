#!/bin/bash
# two-GPU check: slab + two-rank tests, the 2-GPU bench line (with parity), and a fine-profile run
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python tools/show_bench.py gpurun_out/bench_2gpu.json | cut -c1-600
GP_PROF_FINE=1 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 0 --parity-steps 0 > gpurun_out/bench_2gpu_fine.json 2> gpurun_out/bench_2gpu_fine.err
grep "gp fine" gpurun_out/bench_2gpu_fine.err | head -40
