#!/bin/bash
# 2-GPU probe: multi-process tests + bench at 2 ranks
set -u
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout=200 > gpurun_out/gpu_multi_tests.log 2>&1 || echo "MULTI TESTS FAILED"
tail -12 gpurun_out/gpu_multi_tests.log
nvidia-smi --query-compute-apps=pid,used_memory --format=csv
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python tools/show_bench.py gpurun_out/bench_2gpu.json | grep -v n_bodies
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d["slabs"]["parity"])); print(d["slabs"]["per_rank_group_ms"]); print(d["e2e"]); print({k:d["slabs"][k] for k in ("unverified_body_steps","halo_copies_last_step","migrations_last_step")})
P
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29743 bench.py --gpus 2 --steps 20 --warmup 5 --halo 6 --e2e-steps 0 --parity-steps 4 > gpurun_out/bench_2gpu_h6.json 2> gpurun_out/bench_2gpu_h6.err
python tools/show_bench.py gpurun_out/bench_2gpu_h6.json | grep -v n_bodies
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_2gpu_h6.json').read().strip().splitlines()[-1])
print("halo 6:", json.dumps(d["slabs"]["parity"])); print({k:d["slabs"][k] for k in ("unverified_body_steps","halo_copies_last_step","migrations_last_step")})
P
GP_PROF_FINE=1 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 0 --parity-steps 0 > gpurun_out/bench_2gpu_fine.json 2> gpurun_out/bench_2gpu_fine.err
grep "gp fine" gpurun_out/bench_2gpu_fine.err | grep -v "slot" | head -34
