"""CPU: the GEMM form of ActorNet's 1-D convolutions (actor_tc.cu) against torch.nn.functional.conv1d.

The conv engine reads channel-last, zero-padded activations [L + 2][Cin_pad] through overlapping windows: GEMM row p sees
`taps * Cin_pad` consecutive values starting at padded row fold * p * stride (k = 3; one row later for the k = 1 shortcut
conv) and multiplies them with the packed weight [fold * Cout][Kpad].  With fold > 1 one GEMM row produces `fold`
consecutive output steps (columns u * Cout + o), which is how the narrow layers get 64-wide tiles.  The packed operand
comes from the host packer behind mind_debug_conv_fold_pack (no device needed); the window arithmetic is restated here
exactly as actor_tc_run sets up its tensor maps.  Reference layers: planners/mind/networks/layers.py:36-60 (Conv1d,
padding (k - 1) / 2, no bias), network.py:12-61."""
import ctypes as C

import numpy as np
import pytest
import torch


def fold_pack(w, Cin_pad, stride, fold):
    from mind_b200 import lib
    L = lib.load()
    Cout, Cin, ks = w.shape
    w = np.ascontiguousarray(w, dtype=np.float32)
    cap = fold * Cout * (((fold - 1) * stride + ks) * Cin_pad + 64)
    out = np.zeros(cap, np.float32)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    kpad = L.mind_debug_conv_fold_pack(ptr(w), Cout, Cin, Cin_pad, ks, stride, fold, ptr(out), cap)
    assert kpad > 0 and kpad % 64 == 0
    return out[:fold * Cout * kpad].reshape(fold * Cout, kpad), kpad


@pytest.mark.parametrize("Cin,Cin_pad,Cout,ks,stride,fold,L", [
    (14, 16, 32, 3, 1, 2, 48),     # group 0, first conv (input channels padded 14 -> 16)
    (14, 16, 32, 1, 1, 2, 48),     # group 0, shortcut conv
    (32, 32, 32, 3, 1, 2, 48),     # group 0, inner convs
    (32, 32, 32, 3, 1, 4, 48),     # four steps per row
    (32, 32, 64, 3, 2, 1, 48),     # group 1, strided first conv
    (32, 32, 64, 1, 2, 1, 48),     # group 1, strided shortcut
    (32, 32, 64, 3, 2, 2, 48),     # strided and folded
    (64, 64, 128, 3, 1, 1, 24),
])
def test_folded_gemm_equals_conv1d(Cin, Cin_pad, Cout, ks, stride, fold, L):
    rng = np.random.default_rng(Cin * 1000 + Cout + ks + stride * 7 + fold)
    w = rng.normal(0, 0.3, (Cout, Cin, ks)).astype(np.float32)
    x = rng.normal(0, 1, (Cin, L)).astype(np.float32)
    want = torch.nn.functional.conv1d(torch.from_numpy(x)[None].double(), torch.from_numpy(w).double(), stride=stride,
                                      padding=(ks - 1) // 2)[0].numpy()                  # [Cout, Lout]
    Lout = (L - 1) // stride + 1
    assert want.shape == (Cout, Lout) and Lout % fold == 0
    Wf, kpad = fold_pack(w, Cin_pad, stride, fold)
    # channel-last padded activation as the engine stores it, plus the slack the K-padded windows may touch
    xp = np.zeros((L + 2 + 8, Cin_pad), np.float64)
    xp[1:L + 1, :Cin] = x.T
    xp[L + 2:] = rng.normal(0, 1, (8, Cin_pad))          # whatever follows (the next actor): must meet zero weights only
    flat = xp.reshape(-1)
    base = Cin_pad if ks == 1 else 0                      # k = 1: the window starts at the centre row
    got = np.zeros((Cout, Lout))
    for p in range(Lout // fold):
        win = flat[base + p * fold * stride * Cin_pad:][:kpad]
        y = Wf.astype(np.float64) @ win                   # [fold * Cout]
        for u in range(fold):
            got[:, p * fold + u] = y[u * Cout:(u + 1) * Cout]
    assert np.max(np.abs(got - want)) < 1e-9 * max(1.0, np.max(np.abs(want))) + 1e-12
