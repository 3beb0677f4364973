#!/bin/bash
# round evidence: tests, smoke, bench (both arms), ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
R=${1:-r01_final}
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests_$R.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee -a gpurun_out/tests_$R.log
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/bench_$R.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_$R.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 176 -c 176 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_launch_$R.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_rela_fusion_tc -s 6 -c 1 -o gpurun_out/prof_$R -f \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_full_$R.log 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_$R.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'cpu', j['cpu_baseline']['value'], 'tree', j['tree_rollout']['natural']['ms_per_tree'], j['tree_rollout']['forced_full']['ms_per_tree'])"
