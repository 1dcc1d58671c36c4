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
    objdir = os.path.join(os.path.dirname(PKG), "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    base = [nvcc]
    # nvcc's host compiler: the image exports CC/CXX pointing at a gcc without
    # the default specs; use the system g++.
    if os.path.exists("/usr/bin/g++"):
        base += ["-ccbin", "/usr/bin/g++"]
    if verbose:
        base.append("-Xptxas=-v")
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(u):
        obj = os.path.join(objdir, u[:-3] + ".o")
        cmd = base + compile_flags + ["-c", os.path.join(CSRC, u), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, env=dict(os.environ))
        return obj

    # the translation units compile in parallel (the kernel instantiations dominate the build time)
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        objs = list(ex.map(compile_one, units))
    link = base + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + [
        "-Xlinker", "-rpath," + cuda_lib]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link, env=dict(os.environ))
    if os.path.lexists(COMPAT):
        os.remove(COMPAT)
    shutil.copyfile(LIB, COMPAT)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
