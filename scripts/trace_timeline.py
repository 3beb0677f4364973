"""Timeline of the fused layer kernel (development tool): run one 256-scene forward on the -DMIND_TRACE build and print,
per traced tile of CTA 0, the clock64() stamps of the hand-off points relative to the tile's start."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MIND_B200_LIB"] = os.path.join(ROOT, "mind_b200", "libmind_b200_trace.so")
import numpy as np, torch
from mind_b200 import synth, lib as L
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
data = synth.batch_s2(B)
net = ScenePredNetB200(None, dev); net.load_state_dict(sd); net.set_precision("f16tc")
def to_dev(data):
    a, ai, l, li, rpe, tn, tr = data
    return (a.to(dev), [x.to(dev) for x in ai], l.to(dev), [x.to(dev) for x in li],
            [{"scene": r["scene"].to(dev), "scene_mask": None} for r in rpe], tn.to(dev), tr.to(dev))
d = to_dev(data)
for _ in range(2):
    net.forward_packed(d)
torch.cuda.synchronize()
lib = L.load()
n = 8 * 17 * 32
buf = (C.c_longlong * n)()
lib.mind_trace_read.argtypes = [C.c_void_p, C.c_int]
got = lib.mind_trace_read(buf, n)
assert got == n, got
t = np.array(buf[:], dtype=np.int64).reshape(8, 17, 32)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "trace_raw.npy"), t)
names = {0: "loop top", 1: "m1 seen (E1 start)", 2: "E1 pass1 end", 3: "E1 sync1 passed", 4: "E1 pass2 end (arrive a)",
         5: "attend: m2b seen", 6: "attend end (arrive k)", 7: "m2a seen (E2a start)", 8: "passA end", 9: "passB start",
         10: "passB end", 11: "passC start", 12: "passC end (arrive e)"}
inames = {15: "issuer: waits a", 16: "issuer: a seen", 17: "issuer: Gpe issued+commit", 18: "issuer: k seen",
          19: "issuer: GKV issued+commit", 20: "issuer: next tile loaded", 21: "issuer: G1(next) issued", 22: "issuer: e seen"}
for g in range(1, 7):
    base = t[g, :16, 0].min()
    print("tile %d (relative to first warp's loop top; min / median / max over the 16 epilogue warps), tile length %d" %
          (g, t[g + 1, :16, 0].min() - base))
    for k in sorted(names):
        v = t[g, :16, k] - base
        print("   %-28s %6d %6d %6d" % (names[k], v.min(), int(np.median(v)), v.max()))
    for k in sorted(inames):
        print("   %-28s %6d" % (inames[k], t[g, 16, k] - base))
