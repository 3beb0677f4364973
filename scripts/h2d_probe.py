"""Host->device staging probe (development tool): enqueue time vs completion time of 256 x 512 KB pinned copies."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mind_b200 import lib as L
dev = torch.device("cuda", 0)
lib = L.load()
n, m = 256, 160
src = [torch.randn(5, m, m).pin_memory() for _ in range(n)]
big = torch.randn(n, 5, m, m).pin_memory()
s = torch.cuda.Stream(dev)
def run(tag, fn):
    for _ in range(2):
        with torch.cuda.stream(s):
            fn()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-28s enqueue %7.2f ms   complete %7.2f ms   (%.1f GB/s)" % (tag, (t1 - t0) * 1e3, (t2 - t0) * 1e3, n * 5 * m * m * 4 / (t2 - t0) / 1e9))
def per_tensor():
    return [t.to(dev, non_blocking=True) for t in src]
sizes = (C.c_int64 * n)(*[t.numel() * 4 for t in src])
ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in src])
offs = (C.c_int64 * n)()
dst = torch.empty(lib.mind_upload_packed_bytes(sizes, n), dtype=torch.uint8, device=dev)
def packed():
    st = torch.cuda.current_stream(dev).cuda_stream
    assert lib.mind_upload_packed(ptrs, sizes, n, C.c_void_p(dst.data_ptr()), dst.numel(), offs, C.c_void_p(st)) == 0
def one_big():
    return big.to(dev, non_blocking=True)
run("per-tensor .to()", per_tensor)
run("mind_upload_packed", packed)
run("one 131 MB tensor .to()", one_big)
x = torch.empty(64 << 20, device=dev)
def d2h():
    h = getattr(d2h, "h", None)
    if h is None:
        d2h.h = h = torch.empty(x.shape).pin_memory()
    h.copy_(x, non_blocking=True)
n_saved = n
run("D2H 256 MB pinned", d2h)
