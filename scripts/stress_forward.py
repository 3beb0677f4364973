"""Repeat the 256-scene forward and check the kernel-side protocol error flag after every call (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mind_b200 import synth, lib as L
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
data = synth.batch_s2(B)
net = ScenePredNetB200(None, dev); net.load_state_dict(sd); net.set_precision("f16tc")
a, ai, l, li, rpe, tn, tr = data
d = (a.to(dev), [x.to(dev) for x in ai], l.to(dev), [x.to(dev) for x in li],
     [{"scene": r["scene"].to(dev), "scene_mask": None} for r in rpe], tn.to(dev), tr.to(dev))
lib = L.load()
prog = hasattr(lib, "mind_progress_init")
if prog:
    assert lib.mind_progress_init() == 0
ref = None
for it in range(n):
    out = net.forward_packed(d)
    try:
        net.sync_check()
    except Exception as e:
        print("iteration", it, "sync_check:", e, flush=True)
        if prog:
            import ctypes as C
            buf = (C.c_int * (148 * 17))()
            lib.mind_progress_read(buf, 148 * 17)
            import collections
            groups = collections.Counter()
            for cta in range(148):
                w = [buf[cta * 17 + k] for k in range(17)]
                st = [(v >> 8, v & 255) for v in w]
                key = ("issuer %d:%d" % st[0]) + " | epilogue " + " ".join("%d:%d x%d" % (k[0], k[1], n) for k, n in sorted(collections.Counter(st[1:]).items()))
                groups[key] += 1
            for k, n in groups.most_common(12):
                print("%3d CTAs: %s" % (n, k))
        break
    reg = out[1].clone()
    if ref is None:
        ref = reg
    print("iteration", it, "ok, max |reg - reg0| = %.3e" % (reg - ref).abs().max().item(), flush=True)
