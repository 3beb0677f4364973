"""CPU restatement of MIND's ScenePredNet forward.  TEST INFRASTRUCTURE ONLY.

Plain functional fp32 torch on CPU (matmul / elementwise only, no nn.Module),
driven directly by the reference's 328-key state_dict.  Each function cites
the reference lines it restates (paths relative to /root/reference).

Parity pinning: tests/test_oracle_vs_reference.py compares every stage and the
full forward with the reference's own modules (imported through
oracle/ref_loader.py) and tests/test_oracle_golden.py compares with the golden
vectors in tests/golden/ that oracle/make_golden.py dumped from the reference.

Optional `emu` hooks let the numerical-format experiments (fp16 / tf32 operand
rounding in the N^2 contractions) run on CPU before a kernel is written.
"""
import math
from typing import Dict, List, Optional

import numpy as np
import torch

EPS = 1e-5
D = 128
N_MODES = 6
N_PRED = 60
N_ORDER = 7


# --------------------------------------------------------------------------- #
# small building blocks
# --------------------------------------------------------------------------- #
def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x, g, b):
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + EPS) * g + b


def group_norm1(x, g, b):
    """GroupNorm with ONE group over (C, L): planners/mind/networks/layers.py:45-46
    (gcd(ng=1, C) = 1).  x: [A, C, L]."""
    mu = x.mean(dim=(1, 2), keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=(1, 2), keepdim=True)
    return xc / torch.sqrt(var + EPS) * g[None, :, None] + b[None, :, None]


def conv1d(x, w, stride=1):
    """1-D cross-correlation, zero padding (k-1)//2, no bias (layers.py:41-43).
    x [A, Cin, L], w [Cout, Cin, k]."""
    k = w.shape[2]
    pad = (k - 1) // 2
    A, Cin, L = x.shape
    xp = torch.zeros(A, Cin, L + 2 * pad, dtype=x.dtype, device=x.device)
    xp[:, :, pad:pad + L] = x
    Lout = (L + 2 * pad - k) // stride + 1
    out = torch.zeros(A, w.shape[0], Lout, dtype=x.dtype, device=x.device)
    for kk in range(k):
        xs = xp[:, :, kk:kk + stride * (Lout - 1) + 1:stride]  # [A, Cin, Lout]
        out += torch.einsum("oc,acl->aol", w[:, :, kk], xs)
    return out


def upsample_linear_x2(x):
    """F.interpolate(scale_factor=2, mode='linear', align_corners=False)
    (network.py:57).  x [A, C, L] -> [A, C, 2L].
    src = (dst + 0.5)/2 - 0.5 clamped at 0; neighbours clamped to L-1."""
    A, C, L = x.shape
    dst = torch.arange(2 * L, dtype=torch.float32, device=x.device)
    src = torch.clamp((dst + 0.5) * 0.5 - 0.5, min=0.0)
    i0 = src.floor().long()
    i1 = torch.clamp(i0 + 1, max=L - 1)
    lam = (src - i0.float())
    return x[:, :, i0] * (1.0 - lam) + x[:, :, i1] * lam


class Params:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        self.sd = sd
        self.prefix = prefix

    def __call__(self, name):
        return self.sd[self.prefix + name].detach().to(torch.float32)

    def sub(self, p):
        return Params(self.sd, self.prefix + p)

    def has(self, name):
        return (self.prefix + name) in self.sd


# --------------------------------------------------------------------------- #
# ActorNet  (network.py:12-61, layers.py:36-60,140-188)
# --------------------------------------------------------------------------- #
def res1d(x, p: Params, stride):
    out = conv1d(x, p("conv1.weight"), stride)
    out = torch.relu(group_norm1(out, p("bn1.weight"), p("bn1.bias")))
    out = conv1d(out, p("conv2.weight"), 1)
    out = group_norm1(out, p("bn2.weight"), p("bn2.bias"))
    if p.has("downsample.0.weight"):
        x = conv1d(x, p("downsample.0.weight"), stride)
        x = group_norm1(x, p("downsample.1.weight"), p("downsample.1.bias"))
    return torch.relu(out + x)


def actor_net(actors, p: Params):
    """actors [A, 14, 48] -> [A, 128]."""
    out = actors
    outs = []
    for g in range(4):
        out = res1d(out, p.sub("groups.%d.0." % g), 1 if g == 0 else 2)
        out = res1d(out, p.sub("groups.%d.1." % g), 1)
        outs.append(out)

    def lateral(i, x):
        q = p.sub("lateral.%d." % i)
        return group_norm1(conv1d(x, q("conv.weight")), q("norm.weight"), q("norm.bias"))

    out = lateral(3, outs[3])
    for i in (2, 1, 0):
        out = upsample_linear_x2(out) + lateral(i, outs[i])
    out = res1d(out, p.sub("output."), 1)
    return out[:, :, -1]


# --------------------------------------------------------------------------- #
# LaneNet  (network.py:64-121)
# --------------------------------------------------------------------------- #
def _lin_ln_relu(x, p: Params, i):
    return torch.relu(layer_norm(linear(x, p("%d.weight" % i), p("%d.bias" % i)),
                                 p("%d.weight" % (i + 1)), p("%d.bias" % (i + 1))))


def point_aggregate(x_inp, p: Params, aggre_out):
    x = _lin_ln_relu(x_inp, p.sub("fc1."), 0)
    x = _lin_ln_relu(x, p.sub("fc1."), 3)
    m = x.max(dim=1, keepdim=True).values.expand(-1, x.shape[1], -1)
    y = torch.cat([x, m], dim=-1)
    y = _lin_ln_relu(y, p.sub("fc2."), 0)
    y = _lin_ln_relu(y, p.sub("fc2."), 3)
    out = layer_norm(x_inp + y, p("norm.weight"), p("norm.bias"))
    if aggre_out:
        return out.max(dim=1).values
    return out


def lane_net(feats, p: Params):
    """feats [L, 10, 16] -> [L, 128] (never squeezed to 1-D here; the caller
    handles the reference's L==1 squeeze quirk, network.py:97,492-493)."""
    x = _lin_ln_relu(feats, p.sub("proj."), 0)
    x = point_aggregate(x, p.sub("aggre1."), False)
    return point_aggregate(x, p.sub("aggre2."), True)


# --------------------------------------------------------------------------- #
# FusionNet / RelaFusionLayer  (network.py:124-340)
# --------------------------------------------------------------------------- #
def rela_fusion_layer(node, edge, p: Params, update_edge, n_head=8, emu=None):
    """node [N,128], edge [N,N,128] -> (node', edge').

    memory[i,j] = ReLU(LN(W . [edge[i,j]; node[j]; node[i]] + b))   (:197-199)
    edge'       = LN(edge + ReLU(LN(W_pe . memory + b)))             (:202)
    query j attends keys i with k = v = memory[i,j]                   (:222)
    """
    N = node.shape[0]
    r = (lambda t: t) if emu is None else emu
    W = p("proj_memory.0.weight")
    We, Ws, Wt = W[:, :D], W[:, D:2 * D], W[:, 2 * D:]
    S = node @ Ws.t()                       # src_x[i,j] = node[j]
    T = node @ Wt.t() + p("proj_memory.0.bias")   # tar_x[i,j] = node[i]
    pre = r(edge) @ r(We).t() + S[None, :, :] + T[:, None, :]
    memory = torch.relu(layer_norm(pre, p("proj_memory.1.weight"), p("proj_memory.1.bias")))
    if update_edge:
        upd = r(memory) @ r(p("proj_edge.0.weight")).t() + p("proj_edge.0.bias")
        upd = torch.relu(layer_norm(upd, p("proj_edge.1.weight"), p("proj_edge.1.bias")))
        edge = layer_norm(edge + upd, p("norm_edge.weight"), p("norm_edge.bias"))
    Win, bin_ = p("multihead_attn.in_proj_weight"), p("multihead_attn.in_proj_bias")
    q = node @ Win[:D].t() + bin_[:D]                     # [N(j), 128]
    k = r(memory) @ r(Win[D:2 * D]).t() + bin_[D:2 * D]   # [N(i), N(j), 128]
    v = r(memory) @ r(Win[2 * D:]).t() + bin_[2 * D:]
    dh = D // n_head
    qh = q.view(N, n_head, dh) * (1.0 / math.sqrt(dh))
    kh = k.view(N, N, n_head, dh)
    vh = v.view(N, N, n_head, dh)
    s = torch.einsum("jhd,ijhd->jhi", qh, kh)
    a = torch.softmax(s, dim=-1)
    o = torch.einsum("jhi,ijhd->jhd", a, vh).reshape(N, D)
    o = o @ p("multihead_attn.out_proj.weight").t() + p("multihead_attn.out_proj.bias")
    x = layer_norm(node + o, p("norm2.weight"), p("norm2.bias"))
    f = torch.relu(x @ p("linear1.weight").t() + p("linear1.bias"))
    f = f @ p("linear2.weight").t() + p("linear2.bias")
    x = layer_norm(x + f, p("norm3.weight"), p("norm3.bias"))
    return x, edge


def edge_init(rpe, p: Params):
    """rpe [5, M, M] -> edge0 [M+1, M+1, 128] with a zero cls row/col (:326-330)."""
    M = rpe.shape[1]
    e = _lin_ln_relu(rpe.permute(1, 2, 0), p.sub("proj_rpe_scene."), 0)
    out = torch.zeros(M + 1, M + 1, D, device=e.device)
    out[:M, :M] = e
    return out


def fusion_net(actors, actor_idcs, lanes, lane_idcs, rpes, p: Params, emu=None,
               n_layer=6, return_edges=False):
    a = _lin_ln_relu(actors, p.sub("proj_actor."), 0)
    l = _lin_ln_relu(lanes, p.sub("proj_lane."), 0)
    a_new, l_new, c_new, edges = [], [], [], []
    for ai, li, rp in zip(actor_idcs, lane_idcs, rpes):
        x = torch.cat([a[ai], l[li], torch.zeros(1, D, device=a.device)], dim=0)
        edge = edge_init(rp["scene"], p)
        for i in range(n_layer):
            x, edge = rela_fusion_layer(x, edge, p.sub("fuse_scene.fusion.%d." % i),
                                        update_edge=(i != n_layer - 1), emu=emu)
        a_new.append(x[:len(ai)])
        l_new.append(x[len(ai):-1])
        c_new.append(x[-1:])
        edges.append(edge)
    out = (torch.cat(a_new), torch.cat(l_new), torch.cat(c_new))
    return out + (edges,) if return_edges else out


# --------------------------------------------------------------------------- #
# SceneDecoder  (network.py:343-556)
# --------------------------------------------------------------------------- #
def bezier_T(n_order=N_ORDER, n_step=N_PRED):
    ts = np.linspace(0.0, 1.0, n_step, endpoint=True)
    T = [math.comb(n_order, i) * (1.0 - ts) ** (n_order - i) * ts ** i for i in range(n_order + 1)]
    return torch.tensor(np.array(T).T, dtype=torch.float32)          # [60, 8]  (:449-455)


def bezier_Tp(n_order=N_ORDER, n_step=N_PRED):
    ts = np.linspace(0.0, 1.0, n_step, endpoint=True)
    Tp = [n_order * math.comb(n_order - 1, i) * (1.0 - ts) ** (n_order - 1 - i) * ts ** i
          for i in range(n_order)]
    return torch.tensor(np.array(Tp).T, dtype=torch.float32)         # [60, 7]  (:457-464)


def encoder_layer_postnorm(x, p: Params, n_head=4):
    """nn.TransformerEncoderLayer defaults: post-norm, ReLU, batch_first=False.
    x [S, 128] (one scene: sequence = the 6 modes, batch = 1)  (:378-380,502)."""
    S = x.shape[0]
    Win, bin_ = p("self_attn.in_proj_weight"), p("self_attn.in_proj_bias")
    qkv = x @ Win.t() + bin_
    dh = D // n_head
    q = qkv[:, :D].view(S, n_head, dh) * (1.0 / math.sqrt(dh))
    k = qkv[:, D:2 * D].view(S, n_head, dh)
    v = qkv[:, 2 * D:].view(S, n_head, dh)
    a = torch.softmax(torch.einsum("shd,thd->hst", q, k), dim=-1)
    o = torch.einsum("hst,thd->shd", a, v).reshape(S, D)
    o = o @ p("self_attn.out_proj.weight").t() + p("self_attn.out_proj.bias")
    x = layer_norm(x + o, p("norm1.weight"), p("norm1.bias"))
    f = torch.relu(x @ p("linear1.weight").t() + p("linear1.bias"))
    f = f @ p("linear2.weight").t() + p("linear2.bias")
    return layer_norm(x + f, p("norm2.weight"), p("norm2.bias"))


def _mlp2(x, p: Params):
    return _lin_ln_relu(_lin_ln_relu(x, p, 0), p, 3)


def scene_decoder(ctx, actors, actor_idcs, tgt_feat, tgt_rpes, p: Params):
    T, Tp = bezier_T().to(ctx.device), bezier_Tp().to(ctx.device)
    tr = _lin_ln_relu(tgt_rpes, p.sub("proj_rpe."), 0)
    if tgt_feat.dim() == 1:
        tgt_feat = tgt_feat[None]
    tgt = _mlp2(torch.cat([tgt_feat, tr], dim=-1), p.sub("proj_tgt."))
    res_cls, res_reg, res_aux = [], [], []
    for b, ai in enumerate(actor_idcs):
        na = len(ai)
        cls_embed = _mlp2(ctx[b:b + 1], p.sub("ctx_proj.")).view(N_MODES, D)      # [6,128]
        for i in range(2):
            cls_embed = encoder_layer_postnorm(cls_embed, p.sub("ctx_sat.layers.%d." % i))
        actor_embed = _mlp2(actors[ai], p.sub("actor_proj.")).view(na, N_MODES, D).permute(1, 0, 2)
        embed = cls_embed[:, None, :] + actor_embed           # [6, na, 128]
        embed = embed.clone()
        embed[0] = embed[0] + tgt[b][None]                    # quirk: mode 0 of every actor (:506-508)
        c = _mlp2(cls_embed, p.sub("cls."))
        c = linear(c, p("cls.6.weight"), p("cls.6.bias")).view(1, N_MODES)
        c = torch.softmax(c, dim=1)
        h = _mlp2(embed, p.sub("reg."))
        param = linear(h, p("reg.6.weight"), p("reg.6.bias")).view(N_MODES, na, N_ORDER + 1, 5)
        rp = param[..., :2].permute(1, 0, 2, 3)               # [na, 6, 8, 2]
        cp = param[..., 2:].permute(1, 0, 2, 3)               # [na, 6, 8, 3]
        reg = T @ rp
        vel = (Tp @ (rp[:, :, 1:] - rp[:, :, :-1])) / (N_PRED * 0.1)
        cov = T @ cp
        cov_vel = (Tp @ (cp[:, :, 1:] - cp[:, :, :-1])) / (N_PRED * 0.1)
        reg = torch.cat([reg, torch.exp(cov)], dim=-1)        # [na, 6, 60, 5]
        res_cls.append(c)
        res_reg.append(reg)
        res_aux.append((vel, cov_vel, param))
    return res_cls, res_reg, res_aux


# --------------------------------------------------------------------------- #
# full forward  (network.py:582-595)
# --------------------------------------------------------------------------- #
class ScenePredOracle:
    def __init__(self, state_dict, emu=None, device="cpu"):
        """device != "cpu": the same torch-eager restatement with its tensors on that device (bench.py's
        gpu_eager_baseline: what the reference's nn.Module forward amounts to under torch eager on a GPU)."""
        self.device = torch.device(device)
        self.p = Params({k: v.detach().to(self.device).float() for k, v in state_dict.items()})
        self.emu = emu

    @torch.no_grad()
    def stages(self, data):
        actors, actor_idcs, lanes, lane_idcs, rpe, tgt_nodes, tgt_rpe = data
        p = self.p
        a = actor_net(actors.float(), p.sub("actor_net."))
        l = lane_net(lanes.float(), p.sub("lane_net."))
        t = lane_net(tgt_nodes.float(), p.sub("lane_net."))
        a2, l2, c2 = fusion_net(a, actor_idcs, l, lane_idcs, rpe, p.sub("fusion_net."), emu=self.emu)
        return dict(actor_feat=a, lane_feat=l, tgt_feat=t, actors=a2, lanes=l2, cls=c2)

    @torch.no_grad()
    def __call__(self, data):
        actors, actor_idcs, lanes, lane_idcs, rpe, tgt_nodes, tgt_rpe = data
        st = self.stages(data)
        return scene_decoder(st["cls"], st["actors"], actor_idcs, st["tgt_feat"], tgt_rpe.float(),
                             self.p.sub("pred_scene."))


# --------------------------------------------------------------------------- #
# pairwise relative encoding  (planners/mind/utils.py:193-242)
# --------------------------------------------------------------------------- #
def get_rpe(ctrs, vecs, radius=100.0):
    """ctrs, vecs [M,2] -> [5,M,M]; entry [., a, b] relates vecs[b]/ctrs[b] to a:
    cos/sin of angle(vecs[b], vecs[a]), cos/sin of angle(vecs[b], ctrs[b]-ctrs[a]),
    2*|ctrs[b]-ctrs[a]|/radius; eps 1e-10 in the denominators."""
    d = ctrs[None, :, :] - ctrs[:, None, :]            # [a, b] = ctrs[b] - ctrs[a]
    dist = d.norm(dim=-1)
    v1 = vecs[None, :, :].expand(ctrs.shape[0], -1, -1)   # vecs[b]
    v2 = vecs[:, None, :].expand(-1, ctrs.shape[0], -1)   # vecs[a]

    def cs(u, w):
        nu, nw = u.norm(dim=-1), w.norm(dim=-1)
        den = nu * nw + 1e-10
        return ((u[..., 0] * w[..., 0] + u[..., 1] * w[..., 1]) / den,
                (u[..., 0] * w[..., 1] - u[..., 1] * w[..., 0]) / den)

    c1, s1 = cs(v1, v2)
    c2, s2 = cs(v1, d)
    return torch.stack([c1, s1, c2, s2, dist * 2 / radius])
