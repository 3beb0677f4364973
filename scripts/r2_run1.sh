#!/bin/bash
# round 2, GPU call 1: whole GPU suite (no -x, errors printed) + A/B of the prepared fused-kernel variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/r2_ab_product.json | cut -c1-400
for v in best xx xreg2 xreg stpre xreg_stpre p2pre both all3; do
  [ -f mind_b200/libmind_b200_$v.so ] || continue
  echo "== $v"
  MIND_B200_LIB=mind_b200/libmind_b200_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/r2_ab_$v.json | cut -c1-400
done
