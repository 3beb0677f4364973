#!/bin/bash
mkdir -p gpurun_out
R=${1:-v3}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 1 -o gpurun_out/prof_$R -f \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_full_$R.log 2>&1
tail -3 gpurun_out/ncu_full_$R.log
