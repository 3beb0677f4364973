"""Where the host time of pre_process goes (cProfile over 20 calls on pinned inputs, GPU idle and GPU busy)."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
net = ScenePredNetB200(None, dev); net.load_state_dict(bench.load_weights()); net.set_precision("f16tc")
d = bench.make_batch_dict(256, 1000)
def pin(x):
    if isinstance(x, torch.Tensor): return x.pin_memory()
    if isinstance(x, list): return [pin(v) for v in x]
    if isinstance(x, dict): return {k: pin(v) for k, v in x.items()}
    return x
hd = {k: pin(v) for k, v in d.items()}
for mode in (True, False):
    net.rpe_on_device = mode
    staged = net.pre_process(hd); net.forward_packed(staged); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): net.pre_process(hd)
    torch.cuda.synchronize()
    print("rpe_on_device=%s idle GPU: %.2f ms per pre_process" % (mode, (time.perf_counter() - t0) / 20 * 1e3))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        net.forward_packed(staged)          # GPU busy while the next upload is enqueued
        net.pre_process(hd)
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
