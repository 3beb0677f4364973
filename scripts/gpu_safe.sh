#!/bin/bash
# guarded kernel check: every leg under a short timeout (a hung kernel must not eat the GPU budget)
mkdir -p gpurun_out
T=${1:-safe}
timeout 60 python scripts/stress_forward.py 64 6 2>&1 | grep -v "^frame" | tail -1 | cut -c1-200 || echo "stress FAILED/timeout"
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "stress leg failed: stop"; exit 1; }
timeout 90 python scripts/stress_forward.py 256 8 2>&1 | grep -v "^frame" | tail -1 | cut -c1-200
timeout 150 python -m pytest tests/test_forward_gpu.py -q -m gpu -x -k "selftest or tensor_core_stage or f16tc or large" 2>&1 | tail -3
timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/${T}_bench.json | cut -c1-700
