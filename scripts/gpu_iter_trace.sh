#!/bin/bash
mkdir -p gpurun_out
T=${1:-q5}
timeout 90 python scripts/stress_forward.py 256 4 2>&1 | grep -v "^frame" | tail -1 | cut -c1-300 | tee gpurun_out/${T}_stress.log
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/${T}_tests.log
timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/${T}_bench.json | cut -c1-700
[ -f mind_b200/libmind_b200_trace.so ] && timeout 120 python scripts/trace_timeline.py 2>&1 | grep -v "^frame" > gpurun_out/${T}_trace.log
