"""ActorNet at tree-level batch sizes: tensor-core chain (52 launches) against the one-CTA-per-actor SIMT kernel (1 launch)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mind_b200 import synth
from mind_b200.predictor import ScenePredNetB200
dev = torch.device("cuda", 0)
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")
keys = ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"]
for F, na in ((1, 8), (1, 45), (4, 45), (8, 45), (36, 8), (36, 45)):
    row = {"F": F, "actors": F * na}
    for prec in ("fp32", "f16tc"):
        net = ScenePredNetB200(None, dev); net.load_state_dict(sd); net.set_precision(prec)
        data = synth.batch_from_scenes([synth.scene_s1(300 + i, na, 37) for i in range(F)])
        d = net.pre_process(dict(zip(keys, data)))
        for _ in range(3):
            net.forward_packed(d)
        torch.cuda.synchronize()
        net.profile(True); net.profile_read()
        for _ in range(10):
            net.forward_packed(d)
        prof = net.profile_read(); net.profile(False)
        row[prec] = round(prof["actor_net"][0] / 10, 4)
        del net
    print(json.dumps(row))
