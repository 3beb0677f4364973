#!/bin/bash
# quick GPU check of a kernel change: forward / real-scene / single-query parity, protocol stress, kernel-only bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_real_scenes_gpu.py tests/test_zy_single_query_gpu.py tests/test_tree_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep "rel err\|passed\|failed\|Error" | cut -c1-200 > gpurun_out/r2_quick_tests.log
tail -2 gpurun_out/r2_quick_tests.log
timeout 120 python scripts/stress_forward.py 256 4 2>&1 | grep -v "^frame" | tail -1 | cut -c1-200
timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/r2_quick_bench.json | cut -c1-500
