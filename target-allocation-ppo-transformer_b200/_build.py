"""Build the sm_100a shared library (C ABI of include/uavenv_b200.h) in-tree with nvcc.

Output: target-allocation-ppo-transformer_b200/lib/libuavenv_b200.so (git-ignored, travels with gpurun).
nvcc cross-compiles without a GPU, so this also is the repo's "does it build" check
(__graft_entry__.build()).
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libuavenv_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# (source, extra flags).  The env kernels reproduce the reference's fp64 operation order: no FMA contraction.
UNITS = [
    ("uavenv_capi.cu", ["-fmad=false"]),
    ("ppo_gae.cu", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    deps = [os.path.join(ROOT, "include", "uavenv_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            deps.append(os.path.join(CSRC, f))
    return deps


def up_to_date():
    if not os.path.isfile(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    for src, extra in UNITS:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        extra = extra + os.environ.get("UAVENV_EXTRA_NVCC_FLAGS", "").split()
        cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    tmp = LIB_PATH + ".tmp"
    r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", tmp] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
