"""CPU: the closed-form LayerNorm tables of the fp16 edge-init kernel (k_edge_init_ch) against the layer itself.

fusion_net.proj_rpe_scene = Linear(5, 128) + LayerNorm(128) + ReLU (reference planners/mind/networks/network.py:282-286,
applied to the RPE tensor at :326-330).  The kernel never forms the 128 pre-norm channels of a pair row: it evaluates the
variance as a quadratic form of the 5 RPE values and the output as a 7-term product per channel.  The tables come from the
host packer behind mind_debug_edge_init_pack (no device needed); here they are checked against torch's own LayerNorm on
the checkpoint's weights and on random ones, for RPE values over the range get_rpe produces (cos / sin in [-1, 1],
distance * 2 / 100 up to a few units) and for the degenerate all-zero entry of padded rows."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack(W, b, g, be):
    from mind_b200 import lib
    L = lib.load()
    arrs = [np.ascontiguousarray(x, dtype=np.float32) for x in (W, b, g, be)]
    tab, quad = np.zeros(896, np.float32), np.zeros(21, np.float32)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    assert L.mind_debug_edge_init_pack(*[ptr(a) for a in arrs], ptr(tab), ptr(quad)) == 0
    # [32 lanes][2 pairs][7][2] -> [128 channels][7]
    t = tab.reshape(32, 2, 7, 2).transpose(0, 1, 3, 2).reshape(128, 7)
    return t, quad


def closed_form(t, quad, r):
    """what the kernel computes per pair row (fp32 like the device code)"""
    r = r.astype(np.float32)
    var = np.full(r.shape[0], quad[0], np.float32)
    qi = 6
    for k in range(5):
        acc = np.full(r.shape[0], quad[1 + k], np.float32)
        for l in range(k, 5):
            acc = acc + quad[qi] * r[:, l]
            qi += 1
        var = var + r[:, k] * acc
    rstd = (1.0 / np.sqrt(np.maximum(var, 0) + np.float32(1e-5))).astype(np.float32)
    y = (rstd[:, None] * r) @ t[:, :5].T + rstd[:, None] * t[:, 5][None] + t[:, 6][None]
    return np.maximum(y, 0), var


def reference(W, b, g, be, r):
    x = torch.from_numpy(r).double() @ torch.from_numpy(W).double().T + torch.from_numpy(b).double()
    y = torch.nn.functional.layer_norm(x, (128,), torch.from_numpy(g).double(), torch.from_numpy(be).double(), 1e-5)
    return torch.relu(y).numpy(), x.var(dim=1, unbiased=False).numpy()


def rpe_samples(n, seed):
    rng = np.random.default_rng(seed)
    a1, a2 = rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n)
    dist = np.concatenate([rng.uniform(0, 0.2, n // 2), rng.uniform(0, 6.0, n - n // 2)])
    r = np.stack([np.cos(a1), np.sin(a1), np.cos(a2), np.sin(a2), dist], 1)
    r[0] = 0.0                  # padded / degenerate entry
    r[1] = [1, 0, 0, 0, 0]      # a token against itself: cos = 1, everything else 0
    return r.astype(np.float32)


@pytest.mark.parametrize("weights", ["ckpt", "random"])
def test_closed_form_tables_reproduce_the_layer(weights):
    if weights == "ckpt":
        sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")
        W, b = sd["fusion_net.proj_rpe_scene.0.weight"].numpy(), sd["fusion_net.proj_rpe_scene.0.bias"].numpy()
        g, be = sd["fusion_net.proj_rpe_scene.1.weight"].numpy(), sd["fusion_net.proj_rpe_scene.1.bias"].numpy()
    else:
        rng = np.random.default_rng(5)
        W, b = rng.normal(0, 0.6, (128, 5)).astype(np.float32), rng.normal(0.3, 0.5, 128).astype(np.float32)   # non-zero channel mean
        g, be = rng.normal(1, 0.3, 128).astype(np.float32), rng.normal(0, 0.3, 128).astype(np.float32)
    assert W.shape == (128, 5)
    t, quad = pack(W, b, g, be)
    r = rpe_samples(4096, 11)
    got, var = closed_form(t, quad, r)
    want, var_ref = reference(W.astype(np.float32), b, g, be, r)
    assert np.all(var > -1e-6)
    assert np.max(np.abs(var - var_ref) / (var_ref + 1e-5)) < 2e-5
    # the kernel's output is rounded to fp16 (relative 4.9e-4): the closed form must sit far below that
    err = np.max(np.abs(got - want)) / np.max(np.abs(want))
    print("%s weights: closed-form edge init vs LayerNorm, max-norm relative error %.2e" % (weights, err))
    assert err < 5e-6
