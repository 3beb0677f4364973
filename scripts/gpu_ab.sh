#!/bin/bash
# A/B of a variant library against the product: stress + tensor-core tests on the product, kernel-only bench on both
mkdir -p gpurun_out
T=${1:-ab}; V=${2:-nouni}
timeout 90 python scripts/stress_forward.py 256 3 2>&1 | grep -v "^frame" | tail -1 | cut -c1-300 | tee gpurun_out/${T}_stress.log
timeout 200 python -m pytest tests/test_forward_gpu.py tests/test_real_scenes_gpu.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/${T}_tests.log
timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/${T}_bench.json | cut -c1-600
MIND_B200_LIB=mind_b200/libmind_b200_$V.so timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/${T}_bench_$V.json | cut -c1-600
