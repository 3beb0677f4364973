#!/bin/bash
# bench + ncu evidence for the round (1 GPU).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
R=${1:-r01}
echo "== bench f16tc" | tee gpurun_out/bench_$R.log
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee -a gpurun_out/bench_$R.log
echo "== bench fp32 (exact path, smaller batch)" | tee -a gpurun_out/bench_$R.log
timeout 900 python bench.py --steps 3 --warmup 1 --precision fp32 --batch 64 --kernel-only 2>&1 | tail -2 | tee -a gpurun_out/bench_$R.log
echo "== reference arm" | tee -a gpurun_out/bench_$R.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -2 | tee -a gpurun_out/bench_$R.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 1 --kernel-only > gpurun_out/ncu_launch_$R.log 2>&1
echo "== ncu full on the fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 2 -o gpurun_out/prof_$R -f \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_full_$R.log 2>&1
ls -la gpurun_out
