"""Dump golden vectors from the UNMODIFIED reference (run in the build container).

  python -m oracle.make_golden

Writes tests/golden/:
  weights_20240121-172745.pt   state_dict of the shipped checkpoint (328 fp32 tensors;
                               data, not source -- needed because /root/reference does
                               not exist on the GPU box)
  shapes.json                  key -> shape (to regenerate seeded random weights)
  s1_ckpt.npz                  S1 KAT: stage + final outputs, shipped ckpt
  ragged_ckpt.npz              3-scene ragged batch, shipped ckpt
  ragged_rand.npz              same batch, seeded random weights (synth.random_state_dict(0))
Inputs are regenerated from seeds by mind_b200/synth.py, not stored.
"""
import json
import os

import numpy as np
import torch

from oracle import ref_loader
from mind_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_reference(net, data):
    with torch.no_grad():
        a = net.actor_net(data[0])
        l = net.lane_net(data[2])
        t = net.lane_net(data[5])
        a2, l2, c2 = net.fusion_net(a, data[1], l, data[3], data[4])
        cls, reg, aux = net.pred_scene(c2, a2, data[1], t, data[6])
    t2 = t if t.dim() == 2 else t[None]
    d = dict(actor_feat=a, lane_feat=l, tgt_feat=t2, actors=a2, lanes=l2, cls_tok=c2)
    for b in range(len(cls)):
        d["cls_%d" % b] = cls[b]
        d["reg_%d" % b] = reg[b]
        d["vel_%d" % b] = aux[b][0]
        d["covvel_%d" % b] = aux[b][1]
        d["param_%d" % b] = aux[b][2]
    return {k: v.detach().numpy().astype(np.float32) for k, v in d.items()}


def ragged_batch():
    return synth.batch_ragged(batch=3, seed=7, na_rng=(3, 9), nl_rng=(10, 30), seed0=2000)


def main():
    os.makedirs(OUT, exist_ok=True)
    sd = torch.load(ref_loader.CKPT, map_location="cpu")["state_dict"]
    sd = {k: v.detach().float().contiguous() for k, v in sd.items()}
    torch.save(sd, os.path.join(OUT, "weights_20240121-172745.pt"))
    json.dump({k: list(v.shape) for k, v in sd.items()}, open(os.path.join(OUT, "shapes.json"), "w"), indent=0)

    net, _ = ref_loader.build_reference_net(sd)
    np.savez_compressed(os.path.join(OUT, "s1_ckpt.npz"),
                        **run_reference(net, synth.batch_from_scenes([synth.scene_s1(1234)])))
    np.savez_compressed(os.path.join(OUT, "ragged_ckpt.npz"), **run_reference(net, ragged_batch()))
    rsd = synth.random_state_dict(0, like=sd)
    net2, _ = ref_loader.build_reference_net(rsd)
    np.savez_compressed(os.path.join(OUT, "ragged_rand.npz"), **run_reference(net2, ragged_batch()))
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
