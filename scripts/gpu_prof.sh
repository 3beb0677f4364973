#!/bin/bash
# one ncu --set full capture (with source) of the fused layer kernel
mkdir -p gpurun_out
T=${1:-prof}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 1 -o gpurun_out/prof_$T -f \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_full_$T.log 2>&1
tail -3 gpurun_out/ncu_full_$T.log
ls -la gpurun_out/prof_$T.ncu-rep
