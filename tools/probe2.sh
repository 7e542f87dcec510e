#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_multi.py -m gpu -x -q --timeout=200 > gpurun_out/gpu_multi_tests.log 2>&1 || echo "SLAB/MULTI TESTS FAILED"
tail -6 gpurun_out/gpu_multi_tests.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python tools/show_bench.py gpurun_out/bench_2gpu.json | grep -v n_bodies
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d["slabs"]["parity"])); print(d["slabs"]["per_rank_group_ms"]); print(d["e2e"]["value"], d["e2e"]["ms_per_step"]); print({k:d["slabs"][k] for k in ("unverified_body_steps","halo_copies_last_step","migrations_last_step")})
P
GP_PROF_FINE=1 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 0 --parity-steps 0 > gpurun_out/bench_2gpu_fine.json 2> gpurun_out/bench_2gpu_fine.err
grep "slab proof\|k_contacts  " gpurun_out/bench_2gpu_fine.err
