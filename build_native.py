"""In-tree build of libpogs_b200.so (hand-written sm_100a CUDA + the C ABI).

    python build_native.py            # build if sources are newer than the .so
    python build_native.py --force

(kept outside the package so that it can run when the .so is missing or stale: importing
pogs_b200 loads the library and fails loudly without it)

nvcc cross-compiles for sm_100a without a GPU, so this runs on the CPU build box;
the resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pogs_b200")
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libpogs_b200.so")
# Same file under the name the reference's python/pogs/graph.py looks for
# (graph.py:29-67), so the reference wrapper can load it unchanged.
COMPAT = os.path.join(LIBDIR, "libpogs_cpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    out = []
    for root in (CSRC, os.path.join(PKG, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    return out


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "capi.cu"), "-lcublas", "-lcusolver", "-lcusparse",
                                 "-Xlinker", "-rpath," + cuda_lib]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    # nvcc's host compiler: the image exports CC/CXX pointing at a gcc without
    # the default specs; use the system g++.
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, env=env)
    if os.path.lexists(COMPAT):
        os.remove(COMPAT)
    shutil.copyfile(LIB, COMPAT)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
