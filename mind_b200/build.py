"""Build libmind_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
boundary is a plain C ABI loaded through ctypes)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmind_b200.so")
SOURCES = ["mind_api.cu", "simt_kernels.cu", "fusion_tc.cu", "pair_x3.cu", "lane_tc.cu", "node_tc.cu", "tc_gemm.cu", "actor_tc.cu", "tree_step.cu", "cost_field.cu", "ilqr_tree.cpp"]
NVCC_FLAGS = (["-D" + d for d in os.environ.get("MIND_DEFS", "").split()] if os.environ.get("MIND_DEFS") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def source_hash(deps, flags) -> str:
    """sha256 over the compile flags and the CONTENT of every source / header: what decides whether a shipped library
    is the one these sources produce (modification times do not survive a checkout or a snapshot copy)."""
    h = hashlib.sha256(" ".join(flags).encode())
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True builds libmind_b200_trace.so (-DMIND_TRACE: timeline instrumentation of the fused layer kernel,
    a development tool selected with MIND_B200_LIB; never the product library)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    variant = os.environ.get("MIND_VARIANT", "")          # development: -D flags from MIND_DEFS into a separately named library
    LIB = os.path.join(HERE, "libmind_b200_trace.so" if trace else ("libmind_b200_%s.so" % variant if variant else "libmind_b200.so"))
    bdir = os.path.join(HERE, "build_trace" if trace else ("build_" + variant if variant else "build"))
    flags = NVCC_FLAGS + (["-DMIND_TRACE"] if trace else [])
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mind_b200.h"))
    want = source_hash(deps, flags)
    stamp = LIB + ".srchash"
    have = open(stamp).read().strip() if os.path.exists(stamp) else ""
    if not force and os.path.exists(LIB) and have == want:
        return LIB
    force = force or have != want       # a stale or unstamped library is rebuilt from every source
    objs = []
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s) + ".o")
        objs.append(o)
        if force or any(_newer(d, o) for d in deps):
            cmd = [nvcc] + flags + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stderr.write(out)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(want + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
