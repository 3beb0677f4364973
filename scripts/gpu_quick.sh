#!/bin/bash
# quick iteration: tensor-core tests + a kernel-only bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -q -m gpu -s -k "selftest or tensor_core_stage or f16tc or large" 2>&1 | grep -E "passed|failed|PASS|FAIL|Error|rel err|assert [0-9n]" | head -40 | tee gpurun_out/quick.log
timeout 300 python scripts/diag.py 2>&1 | tail -12 | cut -c1-140 | tee -a gpurun_out/quick.log
timeout 300 python bench.py --steps 5 --warmup 2 --kernel-only 2>&1 | tail -2 | tee -a gpurun_out/quick.log
