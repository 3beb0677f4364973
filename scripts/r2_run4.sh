#!/bin/bash
# A/B: W_pe hi-only operand and last-layer G1-ahead (default build has both), each under its own wall-clock guard
mkdir -p gpurun_out
for v in "" _pelo _nog1; do
  lib=mind_b200/libmind_b200$v.so
  [ -f $lib ] || continue
  echo "== variant '$v'"
  MIND_B200_LIB=$lib timeout 120 python scripts/stress_forward.py 256 6 2>&1 | grep -v "^frame" | tail -2 | cut -c1-200
  MIND_B200_LIB=$lib timeout 120 python scripts/stress_forward.py 64 12 2>&1 | grep -v "^frame" | tail -1 | cut -c1-200
  MIND_B200_LIB=$lib timeout 300 python -m pytest tests/test_forward_gpu.py tests/test_zy_single_query_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep "rel err\|passed\|failed" | cut -c1-160 > gpurun_out/r2_ab4_tests$v.log
  tail -1 gpurun_out/r2_ab4_tests$v.log
  MIND_B200_LIB=$lib timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/r2_ab4$v.json | cut -c1-420
done
