#!/bin/bash
# last evidence of the round under a tight budget: ncu full capture of the fused kernel, launch list, then the bench line
mkdir -p gpurun_out
R=${1:-r01_v8}
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 1 -o gpurun_out/prof_$R -f \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_full_$R.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 176 -c 176 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_launch_$R.log 2>&1
timeout 200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_$R.json
cut -c1-600 gpurun_out/bench_$R.json
