#!/bin/bash
timeout 300 python -m pytest tests/test_tree_gpu.py -q -m gpu 2>&1 | tail -2
timeout 300 python - <<'PY'
import torch, json, bench
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
net = ScenePredNetB200(None, dev); net.load_state_dict(bench.load_weights()); net.set_precision("f16tc")
print(json.dumps(bench.bench_tree(net, dev), indent=None))
PY
