#!/bin/bash
# quick iteration: tensor-core tests + a kernel-only bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -q -m gpu -x -k "selftest or tensor_core_stage or f16tc or large" -s 2>&1 | tail -25 | tee gpurun_out/quick.log
timeout 300 python bench.py --steps 5 --warmup 2 --kernel-only 2>&1 | tail -2 | tee -a gpurun_out/quick.log
