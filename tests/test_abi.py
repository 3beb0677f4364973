"""CPU: the C-ABI library builds, loads and exports every symbol include/mind_b200.h declares;
without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_header_symbols():
    from mind_b200 import lib
    L = lib.load()
    hdr = open(os.path.join(ROOT, "include", "mind_b200.h")).read()
    declared = set(re.findall(r"\b(mind_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(L, sym), "library does not export %s" % sym
    assert declared == set(lib.SYMBOLS), (declared ^ set(lib.SYMBOLS))
    assert b"sm_100a" in L.mind_build_info()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from mind_b200 import lib
    from mind_b200.predictor import ScenePredNetB200
    L = lib.load()
    h = C.c_void_p()
    assert L.mind_create(C.byref(h), 0) != 0
    assert b"no CPU fallback" in L.mind_last_error() or b"CUDA" in L.mind_last_error()
    with pytest.raises(RuntimeError):
        ScenePredNetB200(None, torch.device("cpu"))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mind_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_upload_packed_layout_math():
    """mind_upload_packed_bytes: entries are laid out in order at 256-byte aligned offsets (no GPU needed)."""
    import ctypes as C
    from mind_b200 import lib
    L = lib.load()
    sizes = [1, 256, 257, 0, 5 * 160 * 160 * 4]
    arr = (C.c_int64 * len(sizes))(*sizes)
    assert L.mind_upload_packed_bytes(arr, len(sizes)) == 256 + 256 + 512 + 0 + 512000
    assert L.mind_upload_packed_bytes(arr, 0) == 0
    bad = (C.c_int64 * 1)(-4)
    assert L.mind_upload_packed_bytes(bad, 1) == -1
