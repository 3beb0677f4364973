"""Numerical-format experiment (CPU emulation; test infrastructure only).

Question: which operand format can the N^2 contractions of the rela-fusion
layers (W_e, W_pe, W_k, W_v on [N*N,128] rows) use, and can the edge stream
be stored in fp16 between layers, while the final outputs stay within the
1e-3 relative tolerance north_star states?  fp32 accumulation everywhere,
fp32 LayerNorm / softmax / node-side GEMMs.

Run:  python -m oracle.format_experiments      (needs /root/reference for the ckpt)
"""
import sys
import torch

from oracle import ref_loader
from oracle import scene_pred_oracle as O
from mind_b200 import synth


def round_mantissa(x, bits):
    """round-to-nearest-even to `bits` explicit mantissa bits (tf32: 10)."""
    xi = x.contiguous().view(torch.int32)
    drop = 23 - bits
    bias = ((xi >> drop) & 1) + ((1 << (drop - 1)) - 1)
    return (((xi + bias) >> drop) << drop).view(torch.float32)


FORMATS = {
    "fp32": None,
    "tf32": lambda t: round_mantissa(t, 10),
    "fp16": lambda t: t.half().float(),
    "bf16": lambda t: t.bfloat16().float(),
}


def run(sd, data, fmt, edge_store=None):
    emu = FORMATS[fmt]
    orc = O.ScenePredOracle(sd, emu=emu)
    if edge_store is not None:
        # monkeypatch: round the stored edge after every layer
        orig = O.rela_fusion_layer

        def patched(node, edge, p, update_edge, n_head=8, emu=None):
            x, e = orig(node, edge, p, update_edge, n_head, emu)
            return x, FORMATS[edge_store](e)
        O.rela_fusion_layer = patched
        orig_init = O.edge_init
        O.edge_init = lambda rpe, p: FORMATS[edge_store](orig_init(rpe, p))
        try:
            return orc(data)
        finally:
            O.rela_fusion_layer = orig
            O.edge_init = orig_init
    return orc(data)


def main():
    sd = torch.load(ref_loader.CKPT, map_location="cpu")["state_dict"]
    seeds = [1234, 1000, 1001, 1002]
    for seed in seeds:
        data = synth.batch_from_scenes([synth.scene_s1(seed)])
        base = run(sd, data, "fp32")
        print("seed", seed, "cls", [round(float(v), 5) for v in base[0][0][0]])
        for fmt, es in [("tf32", None), ("fp16", None), ("fp16", "fp16"), ("bf16", None), ("bf16", "bf16")]:
            out = run(sd, data, fmt, es)
            dc = (out[0][0] - base[0][0]).abs().max().item()
            rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
            same = torch.equal(out[0][0].argsort(descending=True), base[0][0].argsort(descending=True))
            print("  ops=%-5s edge=%-5s  dcls %.2e  reg_rel %.2e  vel_rel %.2e  order_same %s" %
                  (fmt, es or "fp32", dc, rel(out[1][0], base[1][0]), rel(out[2][0][0], base[2][0][0]), same))


if __name__ == "__main__":
    main()
