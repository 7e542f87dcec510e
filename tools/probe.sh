#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_next_rows.py tests/test_rotation_known_answers.py tests/test_gpu_parity.py -m gpu -x -q --timeout=300 -k "cap_velocity or frames or rotation or instance or aabb or dirty or external" > gpurun_out/gpu_tests.log 2>&1 || echo "TESTS FAILED"
tail -5 gpurun_out/gpu_tests.log
python tools/pcie_probe.py 576 2>&1 | tail -2
GP_PROF_FINE=1 python - > gpurun_out/dirty_probe.log 2>&1 <<'P'
import time, numpy as np, sys
sys.path.insert(0, '.')
import bench
from gears_b200.world import PhysicsWorld
gen, wkw = bench.WORKLOADS["uniform_16M"]
s = bench.make_scene(gen); n = len(s["mass"])
DT = bench.DT
with PhysicsWorld(profile=True, **wkw) as w:
    w.upload(s)
    for _ in range(5): w.step(DT)
    w.kernel_times()
    k = n // 10
    hidx = w.host_alloc((k,), np.uint32); hv = w.host_alloc((k, 3)); ha = w.host_alloc((k, 3))
    hidx[:] = np.sort(np.random.default_rng(1).choice(n, size=k, replace=False)).astype(np.uint32); hv[:] = 0; ha[:] = 0
    w.update_dirty(hidx, vel=hv, acc=ha); w.step_async(DT, 1); w.sync()
    tu = ts = 0.0
    t0 = time.perf_counter()
    for _ in range(10):
        a = time.perf_counter(); w.update_dirty(hidx, vel=hv, acc=ha); b = time.perf_counter(); w.step_async(DT, 1); c = time.perf_counter()
        tu += b - a; ts += c - b
    w.sync(); t1 = time.perf_counter()
    print(f"10% dirty: {1e3*(t1-t0)/10:.2f} ms/step; host time in update_dirty {1e3*tu/10:.2f} ms, in step_async {1e3*ts/10:.2f} ms")
P
grep -v "k_contacts /\|slot " gpurun_out/dirty_probe.log | tail -22
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
python tools/show_bench.py gpurun_out/bench_default.json | grep -v n_bodies
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], "dependent", e["dependent"]["value"], [ (x["dirty_fraction"], x["value"], x["ms_per_step"]) for x in e["dirty_set"]]); print(d["cpu_baseline_grid"]); 
P
CMD="python bench.py --no-cpu --steps 3 --warmup 3 --e2e-steps 2"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_uniform_16M.csv $CMD > gpurun_out/ncu.log 2>&1
python tools/ncu_launches.py gpurun_out/ncu_launches_uniform_16M.csv | head -30
