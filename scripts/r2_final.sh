#!/bin/bash
# final check of a build without the ncu --set full captures: whole GPU suite, smoke, bench (own arm), launch list
mkdir -p gpurun_out
R=${1:-r02_v4}
timeout 900 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/tests_$R.log
tail -3 gpurun_out/tests_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee -a gpurun_out/tests_$R.log
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_$R.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_launch_$R.log 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_$R.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'ms/launch', j['roofline']['ms_per_launch'], 'tree', j['tree_rollout']['natural']['ms_per_tree'], j['tree_rollout']['forced_full']['ms_per_tree']); print(json.dumps(j['stage_ms_per_step']))"
