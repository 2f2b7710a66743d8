"""Builds libagar_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(_HERE, "csrc", "agar_b200.cu")]
DEPS = SOURCES + [
    os.path.join(_HERE, "csrc", "agar_kernels.cuh"),
    os.path.join(_HERE, "csrc", "agar_features.cuh"),
    os.path.join(_HERE, "..", "include", "agar_b200.h"),
    os.path.join(_HERE, "..", "include", "agar_spec.h"),
]
OUT = os.path.join(_HERE, "libagar_b200.so")

# -fmad=false: the engine must reproduce the CPU oracle bit for bit, so no FMA contraction
# (DESIGN.md §3.1); -lineinfo so ncu's source page maps to these files.
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
