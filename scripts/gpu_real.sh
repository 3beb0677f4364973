#!/bin/bash
mkdir -p gpurun_out
T=${1:-real}
timeout 400 python -m pytest tests/test_real_scenes_gpu.py -q -m gpu -s 2>&1 | grep -E "rel err|passed|failed|FAILED|assert|Error" | cut -c1-250 | tee gpurun_out/${T}_real.log
timeout 400 python -m pytest tests -q -m gpu --deselect tests/test_real_scenes_gpu.py 2>&1 | tail -5 | cut -c1-250 | tee gpurun_out/${T}_tests.log
