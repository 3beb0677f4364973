"""CPU: resource budget of the fused layer kernel as ptxas reports it for sm_100a (nvcc cross-compiles without a GPU).
544 threads per CTA leave 96 registers per thread; the epilogue keeps x and Dpe + b in registers across its row-group
barriers, which only pays while the spill traffic stays small."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(shutil.which(NVCC) is None, reason="nvcc not available")
def test_fused_layer_kernel_resources(tmp_path):
    src = os.path.join(ROOT, "mind_b200", "csrc", "fusion_tc.cu")
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v",
           "-I" + os.path.join(ROOT, "mind_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-cubin", "-o", str(tmp_path / "v.cubin"), src]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"k_rela_fusion_tc.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers",
                  out.stderr, re.S)
    assert m, out.stderr[-1500:]
    stack, st, ld, regs = (int(v) for v in m.groups())
    print("k_rela_fusion_tc: %d registers, %d B stack, %d / %d B spill stores / loads" % (regs, stack, st, ld))
    assert regs <= 96
    assert st <= 128 and ld <= 128
