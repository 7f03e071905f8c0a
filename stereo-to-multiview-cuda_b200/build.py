"""In-tree build of libs2mv.so (explicit nvcc, sm_100a only)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libs2mv.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: every fused multiply-add in the kernels is an explicit __fmaf_rn
# placed where the reference's PTX has one; nothing else may be contracted.
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]
SOURCES = ["s2mv_api.cu", "s2mv_compat.cu"]


def _sources_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".inl", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _sources_mtime():
        return LIB_PATH
    if not os.path.exists(NVCC):
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise RuntimeError("nvcc not found and libs2mv.so not prebuilt")
    cmd = [NVCC, *NVCC_FLAGS, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
