#!/bin/bash
# First-contact GPU run: self test, exact-path parity, tensor-core parity (separate processes so a
# trapped kernel in one leg cannot mask the others).  Everything is bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== selftest" | tee gpurun_out/check.log
timeout 300 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "selftest" -s 2>&1 | tail -15 | tee -a gpurun_out/check.log
echo "== fp32 path" | tee -a gpurun_out/check.log
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -k "fp32 or writable" -s 2>&1 | tail -30 | tee -a gpurun_out/check.log
echo "== tc path" | tee -a gpurun_out/check.log
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -k "f16tc or large" -s 2>&1 | tail -40 | tee -a gpurun_out/check.log
