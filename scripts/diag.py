import sys, json, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mind_b200 import synth
from mind_b200.predictor import ScenePredNetB200
from oracle import scene_pred_oracle as O
dev = torch.device("cuda", 0)
ck = torch.load("tests/golden/weights_20240121-172745.pt")
rd = synth.random_state_dict(0, like=json.load(open("tests/golden/shapes.json")))
def to_dev(data):
    a, ai, l, li, rpe, tn, tr = data
    return (a.to(dev), [x.to(dev) for x in ai], l.to(dev), [x.to(dev) for x in li],
            [{"scene": r["scene"].to(dev), "scene_mask": None} for r in rpe], tn.to(dev), tr.to(dev))
def rel(a, b): return ((a.cpu() - b).abs().max() / b.abs().max()).item()
for wname, sd in (("ckpt", ck), ("rand", rd)):
    p = O.Params({k: v.float() for k, v in sd.items()})
    for na_list in ([32], [18], [7, 11], [32] * 6, [5, 64, 3], [64] * 4):
        scenes = [synth.scene_s1(900 + i, na, 12) for i, na in enumerate(na_list)]
        data = synth.batch_from_scenes(scenes)
        ref_actor = O.actor_net(data[0], p.sub("actor_net."))
        ref = O.ScenePredOracle(sd)(data)
        out = {}
        for mode, simt in (("tc-actor", 0), ("simt-actor", 1)):
            net = ScenePredNetB200(None, dev); net.load_state_dict(sd); net.set_precision("f16tc"); net.set_option("actor_simt", simt)
            pk = net.forward_packed(to_dev(data))
            af = net.debug_tap("actor_feat", data[0].shape[0] * 128).view(-1, 128)
            torch.cuda.synchronize()
            out[mode] = (rel(af, ref_actor), rel(pk[1], torch.cat(ref[1])))
        print(wname, na_list[:3], len(na_list), "actor_feat tc %.2e simt %.2e | reg err tc-actor %.2e simt-actor %.2e" %
              (out["tc-actor"][0], out["simt-actor"][0], out["tc-actor"][1], out["simt-actor"][1]), flush=True)
