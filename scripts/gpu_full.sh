#!/bin/bash
# round-end style check: gpu tests, smoke, bench (both arms)
mkdir -p gpurun_out
R=${1:-r01b}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/full_$R.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/full_$R.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_$R.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_$R.json | cut -c1-600
