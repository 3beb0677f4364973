"""Where does the host time of the e2e serving loop go? (development tool)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
net = ScenePredNetB200(None, dev); net.load_state_dict(bench.load_weights()); net.set_precision("f16tc")
B = 256
host = bench.make_batch(B, 1000)
keys = ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"]
def pin(x):
    if isinstance(x, torch.Tensor): return x.pin_memory()
    if isinstance(x, list): return [pin(v) for v in x]
    if isinstance(x, dict): return {k: pin(v) for k, v in x.items()}
    return x
hd = {k: pin(v) for k, v in zip(keys, host)}
copy_s = torch.cuda.Stream(dev); comp_s = torch.cuda.current_stream(dev)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()
for mode in ("idle", "busy"):
    T.clear()
    for it in range(6):
        if mode == "busy":
            net.forward_packed(net.pre_process(hd))      # GPU busy with a forward while we time the next enqueue
        t = time.perf_counter()
        with torch.cuda.stream(copy_s):
            a = hd["ACTORS"].to(dev, non_blocking=True); t = tick("to(ACTORS)", t)
            l = hd["LANES"].to(dev, non_blocking=True); t = tick("to(LANES)", t)
            r = net._upload_rpe(hd["RPE"]); t = tick("upload_rpe", t)
            tn = hd["TGT_NODES"].to(dev, non_blocking=True); tr = hd["TGT_RPE"].to(dev, non_blocking=True); t = tick("to(TGT)", t)
        data = (a, hd["ACTOR_IDCS"], l, hd["LANE_IDCS"], r, tn, tr)
        ev = torch.cuda.Event(); ev.record(copy_s); comp_s.wait_event(ev); t = tick("event", t)
        pk = net.forward_packed(data); t = tick("forward_packed", t)
        out = net.forward(data); t = tick("forward(+split)", t)
        torch.cuda.synchronize(); t = tick("sync", t)
    print(mode, {k: round(v / 6 * 1e3, 3) for k, v in T.items()})
