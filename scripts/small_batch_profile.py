"""Stage split of the forward at tree-level batch sizes (F scenes of na actors x nl lanes), both tiers of the
tensor-core mode, plain launches with the library's CUDA-event ranges and graph replay device time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mind_b200 import synth
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")
keys = ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"]
na, nl = int(sys.argv[1]) if len(sys.argv) > 1 else 45, int(sys.argv[2]) if len(sys.argv) > 2 else 37
for min_tok in (128, 0):
    for F in (1, 4, 36, 216):
        net = ScenePredNetB200(None, dev); net.load_state_dict(sd); net.set_precision("f16tc")
        net.set_option("tc_min_tokens", min_tok)
        data = synth.batch_from_scenes([synth.scene_s1(300 + i, na, nl) for i in range(F)])
        d = net.pre_process(dict(zip(keys, data)))
        for _ in range(3):
            net.forward_packed(d)
        torch.cuda.synchronize()
        net.profile(True); net.profile_read()
        reps = 10
        for _ in range(reps):
            net.forward_packed(d)
        prof = net.profile_read(); net.profile(False)
        l0 = net.launch_count(); net.forward_packed(d); l1 = net.launch_count()
        net.use_graphs(True)
        for _ in range(4):
            net.forward_packed(d, persistent_out=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            net.forward_packed(d, persistent_out=True)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"tier": "exact" if min_tok else "fused", "F": F, "tokens": na + nl + 1, "launches": l1 - l0,
                          "graph_ms": round(e0.elapsed_time(e1) / 20, 4),
                          "stage_ms": {k: round(v[0] / reps, 4) for k, v in sorted(prof.items())}}))
        del net
