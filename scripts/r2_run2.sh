#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/r2_tests2.log
cat gpurun_out/r2_tests2.log
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench2.json
cut -c1-600 gpurun_out/r2_bench2.json
