#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | grep "plan calls\|passed\|failed\|Error" | cut -c1-420 > gpurun_out/r2_tests7.log
cat gpurun_out/r2_tests7.log
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench7.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench7.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['tree_rollout'].items():
    if isinstance(v,dict): print(k, round(v.get('ms_per_tree'),2), v.get('host_phase_ms_last'))
PY
