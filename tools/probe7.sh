#!/bin/bash
# repeat the slab tests (the proof's unions now run beside the sweep's tail levels): any flakiness shows here
set -u
mkdir -p gpurun_out
for i in 1 2 3 4; do
timeout -k 10 300 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1
done
