#!/bin/bash
# round-2 evidence: tests, smoke, bench (both arms), ncu launch list, ncu --set full of the dominant kernel and of the
# other kernels of the step (GEMM engine, edge init, GroupNorm apply, cost fields)
mkdir -p gpurun_out
R=${1:-r02}
timeout 900 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/tests_$R.log
tail -3 gpurun_out/tests_$R.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee -a gpurun_out/tests_$R.log
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_$R.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_launch_$R.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_rela_fusion_tc -s 6 -c 1 -o gpurun_out/prof_$R python bench.py --steps 1 --warmup 1 --kernel-only > gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_rela_fusion_tc -s 11 -c 1 -o gpurun_out/prof_last_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_tc_gemm -s 47 -c 1 -o gpurun_out/prof_gemm_actor_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_lane_net_tc -s 1 -c 1 -o gpurun_out/prof_lane_chain_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_node_chain_tc -s 7 -c 1 -o gpurun_out/prof_node_chain_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_edge_init_ch -s 1 -c 1 -o gpurun_out/prof_edge_init_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_gn_apply -s 16 -c 1 -o gpurun_out/prof_gn_apply_$R python bench.py --steps 1 --warmup 1 --kernel-only >> gpurun_out/ncu_full_$R.log 2>&1
timeout 300 $NCU -k regex:k_node_fields -s 1 -c 1 -o gpurun_out/prof_node_fields_$R python -m pytest tests/test_zz_cost_field_gpu.py -q -m gpu -p no:cacheprovider >> gpurun_out/ncu_full_$R.log 2>&1
ls -la gpurun_out/*_$R*
python -c "
import json; j=json.load(open('gpurun_out/bench_$R.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'cpu', j['cpu_baseline']['value'], 'tree', j['tree_rollout']['natural']['ms_per_tree'], j['tree_rollout']['forced_full']['ms_per_tree'])"
