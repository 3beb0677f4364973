#!/bin/bash
# re-entry check: full gpu tests on HEAD (key-split build), kernel-only bench, FFMA2 microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2a_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/r2a_bench.json | cut -c1-1800
timeout 60 scripts/mb/bin/ffma2 2>&1 | tee gpurun_out/r2a_ffma2.log
