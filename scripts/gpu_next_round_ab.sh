#!/bin/bash
# First GPU call of the next round: A/B of the prepared fused-kernel variants against the product library.
#   build here (CPU) first:   for v in best xx xreg2 xreg stpre xreg_stpre p2pre both all3; do ... ; done   (see the loop below), then
#   gpurun --timeout 500 -- 'bash scripts/gpu_next_round_ab.sh'
# Variants (macro-guarded in csrc/fusion_tc.cu, default build unchanged; profiles/r01_v8_stall_analysis.md):
#   stpre  -DMIND_EXP_STPRE   S[j] + T[i] formed in the shadow of the wait for G1 (E1 pass 1)
#   p2pre  -DMIND_EXP_P2PRE   LN_mem gamma / beta of the first 16 channels loaded in front of the row-group barrier (E1 pass 2)
#   xreg   -DMIND_EXP_XREG    x kept in 32 registers across the row-group barrier instead of a TMEM scratch round trip
#                             (fewer spills than the default build: 76 / 72 B against 100 / 88 B)
#   xreg2  -DMIND_EXP_XREG2   Dpe + b kept in 32 registers from edge pass A to pass C (no TMEM scratch round trips, b_pe loaded once)
#   xreg_stpre, both (stpre + p2pre), all3 (xreg + stpre + p2pre), xx (xreg + xreg2), best (xreg + xreg2 + stpre:
#   28 / 36 B of spills against 100 / 88 B, LDTM 20 -> 14, STTM 10 -> 4, LDS 140 -> 132 static)
if [ "$1" = "build" ]; then
  MIND_VARIANT=stpre MIND_DEFS="MIND_EXP_STPRE" python mind_b200/build.py
  MIND_VARIANT=p2pre MIND_DEFS="MIND_EXP_P2PRE" python mind_b200/build.py
  MIND_VARIANT=xreg MIND_DEFS="MIND_EXP_XREG" python mind_b200/build.py
  MIND_VARIANT=xreg_stpre MIND_DEFS="MIND_EXP_XREG MIND_EXP_STPRE" python mind_b200/build.py
  MIND_VARIANT=both MIND_DEFS="MIND_EXP_STPRE MIND_EXP_P2PRE" python mind_b200/build.py
  MIND_VARIANT=all3 MIND_DEFS="MIND_EXP_XREG MIND_EXP_STPRE MIND_EXP_P2PRE" python mind_b200/build.py
  MIND_VARIANT=xreg2 MIND_DEFS="MIND_EXP_XREG2" python mind_b200/build.py
  MIND_VARIANT=xx MIND_DEFS="MIND_EXP_XREG MIND_EXP_XREG2" python mind_b200/build.py
  MIND_VARIANT=best MIND_DEFS="MIND_EXP_XREG MIND_EXP_XREG2 MIND_EXP_STPRE" python mind_b200/build.py
  exit 0
fi
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/ab_product.json | cut -c1-500
for v in best xx xreg2 xreg stpre xreg_stpre p2pre both all3; do
  [ -f mind_b200/libmind_b200_$v.so ] || continue
  MIND_B200_LIB=mind_b200/libmind_b200_$v.so timeout 90 python scripts/stress_forward.py 256 3 2>&1 | grep -v "^frame" | tail -1 | cut -c1-200
  MIND_B200_LIB=mind_b200/libmind_b200_$v.so timeout 200 python -m pytest tests/test_forward_gpu.py tests/test_real_scenes_gpu.py -q -m gpu -x 2>&1 | tail -2 | cut -c1-200 | tee gpurun_out/ab_tests_$v.log
  MIND_B200_LIB=mind_b200/libmind_b200_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --kernel-only 2>&1 | tail -1 | tee gpurun_out/ab_$v.json | cut -c1-500
done
