"""Forward at tree-level batch sizes (F scenes of 8 actors x 60 lanes): wall time per call and, under ncu, the launch list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mind_b200 import synth
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")
net = ScenePredNetB200(None, dev); net.load_state_dict(sd); net.set_precision("f16tc")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
data = synth.batch_from_scenes([synth.scene_s1(300 + i, 8, 60) for i in range(F)])
keys = ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"]
d = net.pre_process(dict(zip(keys, data)))
graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
if graph:
    net.use_graphs(True)
_fp = net.forward_packed
net.forward_packed = (lambda x: _fp(x, persistent_out=True)) if graph else _fp
ref = [t.clone() for t in _fp(d)[:3]]
for _ in range(3):
    net.forward_packed(d)
got = net.forward_packed(d)[:3]
torch.cuda.synchronize()
print("graph" if graph else "plain", "max |diff| vs first plain forward:", [float((a - b).abs().max()) for a, b in zip(got, ref)])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(reps):
    net.forward_packed(d)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize()
l0 = net.launch_count(); net.forward_packed(d); l1 = net.launch_count()
print("graph replays", net.graph_replays())
print("F=%d: device %.3f ms / forward, host enqueue %.3f ms / forward, %d launches" % (F, e0.elapsed_time(e1) / reps, (t1 - t0) * 1e3 / reps, l1 - l0))
