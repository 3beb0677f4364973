#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/r2_tests3.log
tail -4 gpurun_out/r2_tests3.log
grep -n "FAIL\|Error\|error" gpurun_out/r2_tests3.log | head -20
timeout 300 python scripts/small_batch_profile.py 45 37 > gpurun_out/r2_small3.log 2>&1
cat gpurun_out/r2_small3.log | cut -c1-600
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench3.json
cut -c1-300 gpurun_out/r2_bench3.json
