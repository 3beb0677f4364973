#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU); full log kept, every leg under timeout
N=${1:-2}; TAG=${2:-r02_v3}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.log 2>&1
echo "rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_$TAG.log | tail -1 > gpurun_out/bench_${N}gpu_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${N}gpu_$TAG.json"))
print({k: d.get(k) for k in ["value","n_gpus","ms_per_step","scaling"]}, "e2e", d.get("e2e",{}).get("value"))
print("sharded tree:", json.dumps(d.get("tree_rollout_sharded"))[:600])
PY
tail -3 gpurun_out/bench_${N}gpu_$TAG.log | cut -c1-300
