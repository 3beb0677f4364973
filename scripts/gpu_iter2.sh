#!/bin/bash
# kernel iteration: tensor-core parity tests + kernel-only bench + timeline trace
mkdir -p gpurun_out
T=${1:-iter}
timeout 600 python -m pytest tests/test_forward_gpu.py -q -m gpu -x -k "selftest or tensor_core_stage or f16tc or large" 2>&1 | tail -4 | tee gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/${T}_bench.json | cut -c1-900
timeout 300 python scripts/trace_timeline.py > gpurun_out/${T}_trace.log 2>&1; sed -n '/^tile 4/,/^tile 5/p' gpurun_out/${T}_trace.log
