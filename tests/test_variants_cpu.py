"""CPU: the macro-guarded A/B variants of the fused layer kernel prepared for the next GPU session
(scripts/gpu_next_round_ab.sh, profiles/r01_v8_stall_analysis.md) keep compiling for sm_100a, and the combination meant to
be tried first does not spill more than the product build.  nvcc cross-compiles without a GPU; one compile ~10 s."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def ptxas_spills(defs, tmp_path):
    src = os.path.join(ROOT, "mind_b200", "csrc", "fusion_tc.cu")
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v",
           "-I" + os.path.join(ROOT, "mind_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-cubin", "-o", str(tmp_path / "v.cubin"), src]
    out = subprocess.run(cmd + ["-D" + d for d in defs], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"k_rela_fusion_tc.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers",
                  out.stderr, re.S)
    assert m, out.stderr[-1500:]
    return tuple(int(v) for v in m.groups())


@pytest.mark.skipif(shutil.which(NVCC) is None, reason="nvcc not available")
def test_prepared_variants_compile(tmp_path):
    base = ptxas_spills([], tmp_path)
    best = ptxas_spills(["MIND_EXP_XREG", "MIND_EXP_XREG2", "MIND_EXP_STPRE"], tmp_path)
    assert base[3] == 96 and best[3] <= 96, (base, best)                 # 544 threads: 96 registers is the ceiling
    assert best[1] <= base[1] and best[2] <= base[2], (base, best)      # fewer spill bytes than the product build
    p2 = ptxas_spills(["MIND_EXP_P2PRE"], tmp_path)
    assert p2[3] <= 96
