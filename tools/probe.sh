#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout -k 10 600 python bench.py --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; }
run base_16M X=1 -- --workload uniform_16M --steps 20 --warmup 5
python - <<'P'
import json
d=json.loads(open('gpurun_out/base_16M.json').read().strip().splitlines()[-1])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "dependent", e["dependent"]["value"])
for x in e["dirty_set"]: print(x["dirty_fraction"], x["value"], x["ms_per_step"], x["step_groups_ms"], sum(x["step_groups_ms"].values()))
P
for v in inner035:GP_INNER_FRAC=0.35 inner018:GP_INNER_FRAC=0.18 occ025:GP_TARGET_OCC=0.25 occ04:GP_TARGET_OCC=0.4; do
  run ${v%%:*}_16M ${v##*:} -- --workload uniform_16M --steps 20 --warmup 5 --e2e-steps 0
  run ${v%%:*}_2M ${v##*:} -- --workload uniform_2M --steps 20 --warmup 5 --e2e-steps 0
done
run base_2M X=1 -- --workload uniform_2M --steps 20 --warmup 5 --e2e-steps 0
for t in 3 4; do
  GP_NVCC_EXTRA="-DGP_WP_PER_LANE=$t" python -c "from gears_b200 import build; build.build(force=True)" > gpurun_out/build_pl$t.log 2>&1
  run pl${t}_16M X=1 -- --workload uniform_16M --steps 20 --warmup 5 --e2e-steps 0
  run pl${t}_2M X=1 -- --workload uniform_2M --steps 20 --warmup 5 --e2e-steps 0
done
python -c "from gears_b200 import build; build.build(force=True)" > /dev/null 2>&1
run pile_settled X=1 -- --workload pile_1M --steps 100 --warmup 300 --e2e-steps 0
run mixed_4M X=1 -- --workload mixed_4M --steps 20 --warmup 5
python tools/show_bench.py gpurun_out/base_*.json gpurun_out/inner*.json gpurun_out/occ0*.json gpurun_out/pl*.json gpurun_out/pile_settled.json gpurun_out/mixed_4M.json | grep -v "n_bodies\|e2e"
