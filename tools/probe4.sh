#!/bin/bash
# variant sweep: every build_variants/*.so takes the library's place for one kernel-only bench line per workload
set -u
mkdir -p gpurun_out
cp gears_b200/libgears_physics_cuda.so /tmp/default.so
for so in build_variants/*.so; do
v=$(basename $so .so)
cp $so gears_b200/libgears_physics_cuda.so
python -m pytest tests/test_gpu_primitives.py -q -m gpu -x 2>&1 | tail -1
for wl in ${WORKLOADS:-uniform_16M uniform_2M}; do
python bench.py --workload $wl --no-cpu --e2e-steps 0 > gpurun_out/v_${v}_$wl.json 2> gpurun_out/v_${v}_$wl.err
echo "$v $wl"; python tools/show_bench.py gpurun_out/v_${v}_$wl.json | grep -v n_bodies | cut -c1-420
done
done
cp /tmp/default.so gears_b200/libgears_physics_cuda.so
