#!/bin/bash
set -u
mkdir -p gpurun_out
for n in ${NS:-8 4}; do
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2975$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
tail -2 gpurun_out/bench_${n}gpu.err
python tools/show_bench.py gpurun_out/bench_${n}gpu.json 2>/dev/null | head -2 | cut -c1-420
python - <<P
import json
d=json.loads(open('gpurun_out/bench_${n}gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d["slabs"]["parity"])); 
for r in d["slabs"]["per_rank_group_ms"]: print(r)
print(d["e2e"]["value"], d["e2e"]["ms_per_step"]); print({k:d["slabs"][k] for k in ("unverified_body_steps","halo_copies_last_step","migrations_last_step")})
P
done
if [ "${MIXED:-1}" = 1 ]; then
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29763 bench.py --gpus 8 --steps 20 --warmup 5 --workload mixed_4M > gpurun_out/bench_mixed_8gpu.json 2> gpurun_out/bench_mixed_8gpu.err
tail -2 gpurun_out/bench_mixed_8gpu.err
python tools/show_bench.py gpurun_out/bench_mixed_8gpu.json 2>/dev/null | head -2 | cut -c1-420
python - <<P
import json
d=json.loads(open('gpurun_out/bench_mixed_8gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d["slabs"]["parity"])); print({k:d["slabs"][k] for k in ("unverified_body_steps","halo_copies_last_step","migrations_last_step")})
P
fi
[ "${FINE:-1}" = 1 ] || exit 0
GP_PROF_FINE=1 timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 0 --parity-steps 0 > gpurun_out/bench_8gpu_fine.json 2> gpurun_out/bench_8gpu_fine.err
grep "gp fine" gpurun_out/bench_8gpu_fine.err | awk '/us\/step/ {for(i=1;i<=NF;i++) if($i=="us/step") {v=$(i-1); name=""; for(j=3;j<i-1;j++) name=name" "$j; s[name]+=v; c[name]++}} END {for (n in s) printf "%-45s %8.1f us/step (mean of %d ranks)\n", n, s[n]/c[n], c[n]}' | sort | grep -v slot
