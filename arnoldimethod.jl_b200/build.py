"""Build recipe for libb200arnoldi.so (sm_100a only, in-tree).

    python arnoldimethod.jl_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The shared object lands next to this file
so that it travels with the repo snapshot to the GPU box.
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200arnoldi.so")

SOURCES = [os.path.join(CSRC, "b2a.cu")]
HEADERS = [
    os.path.join(ROOT, "include", "b200arnoldi.h"),
    *[os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".hpp"))],
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
    "-ldl",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False, ptxas_info=False):
    if not force and not needs_build():
        return LIB
    tmp = LIB + ".building"  # build beside, then rename: a snapshot of the tree never sees a half-written library
    cmd = [_nvcc(), *NVCC_FLAGS, *SOURCES, "-o", tmp]
    if ptxas_info:
        cmd[1:1] = ["-Xptxas", "-v"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed building libb200arnoldi.so")
    os.replace(tmp, LIB)
    if verbose or ptxas_info:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))
