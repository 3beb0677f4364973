#!/bin/bash
# ncu --set full of three k_tc_gemm instances of the final build: lateral 0 (GroupNorm + FPN epilogue, r_in = 64), group-0 conv2
# (two GroupNorms + ReLU epilogue, folded rows) and the decoder's reg.0 linear (plain epilogue, 49k rows)
R=${1:-r02_v4}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:k_tc_gemm -s 55 -c 1 -o gpurun_out/prof_gemm_lat0_$R python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_gemm_$R.log 2>&1
timeout 200 $NCU -k regex:k_tc_gemm -s 34 -c 1 -o gpurun_out/prof_gemm_g0c2_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_gemm_$R.log 2>&1
timeout 200 $NCU -k regex:k_tc_gemm -s 61 -c 1 -o gpurun_out/prof_gemm_reg0_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_gemm_$R.log 2>&1
ls -la gpurun_out/prof_gemm_*_$R.ncu-rep
