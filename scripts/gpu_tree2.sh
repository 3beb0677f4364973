#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tree_gpu.py -q -m gpu -x 2>&1 | tail -4
for f in 1 4 36; do timeout 100 python scripts/tiny_forward.py $f 20 2>&1 | tail -2; timeout 100 python scripts/tiny_forward.py $f 20 graph 2>&1 | tail -2; done
timeout 300 python - <<'EOF'
import json, torch, bench
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
net = ScenePredNetB200(None, dev); net.load_state_dict(bench.load_weights()); net.set_precision("f16tc")
print(json.dumps(bench.bench_tree(net, dev)))
EOF
