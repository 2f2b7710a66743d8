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
    os.path.join(_HERE, "csrc", "agar_host_expand.h"),
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


def build_specialised(cfg, verbose=False):
    """Compiles a copy of the library whose step kernel is specialised at compile time for `cfg`'s sizes (the
    "config view" of csrc/agar_kernels.cuh: same source, record offsets and loop bounds as constants; typically
    25-35 % faster than the run-time-sized kernel).  Needs nvcc on the machine, takes ~30 s once per distinct
    (agents, bots, pellets, viruses, foods, grid, channels, ticks) tuple; the result is cached in-tree under
    agarai_b200/_jit/.  Returns the path; pass it to BatchedAgarioEnv(..., lib_path=...) or use make(...,
    specialise=True).  Configs with an event log keep the generic kernel (the specialised ones compile events out)."""
    import hashlib
    key = (cfg.num_agents, cfg.num_bots, cfg.num_pellets, cfg.num_viruses, cfg.max_foods, cfg.grid_size, cfg.obs_flags,
           cfg.ticks_per_step)
    names = ("A", "BOTS", "P", "V", "F", "G", "FLAGS", "TICKS")
    # CTAs per SM the shared-memory footprint allows (host-side replica of the bound agar_create computes): the
    # generic launch-bounds ladder 5..8
    src_hash = hashlib.sha1(b"".join(open(d, "rb").read() for d in DEPS) + " ".join(NVCC_FLAGS).encode()).hexdigest()[:10]
    out_dir = os.path.join(_HERE, "_jit")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libagar_b200_%s_%s.so" % ("_".join(str(k) for k in key), src_hash))
    if os.path.exists(out):
        return out
    blocks = _min_blocks_for(cfg)
    tmp = "%s.%d.tmp" % (out, os.getpid())  # one process per GPU may build the same key at once: private file, atomic rename
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-DAGAR_EXTRA_CV", "-DAGAR_EXTRA_CV_BLOCKS=%d" % blocks] + ["-DAGAR_EXTRA_%s=%d" % nv for nv in zip(names, key)] + [
        "-o", tmp] + SOURCES
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    # builds of the same size tuple from older sources (another hash) are dead weight: drop them
    prefix = "libagar_b200_%s_" % "_".join(str(k) for k in key)
    for name in os.listdir(out_dir):
        if name.startswith(prefix) and name.endswith(".so") and os.path.join(out_dir, name) != out:
            try:
                os.remove(os.path.join(out_dir, name))
            except OSError:
                pass
    return out


def _min_blocks_for(cfg):
    """CTAs per SM by shared-memory footprint, through the built library's own layout arithmetic."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    fn = L.agar_step_smem_bytes
    fn.restype = C.c_int64
    fn.argtypes = [C.POINTER(_lib.AgarConfig)]
    smem = int(fn(C.byref(cfg)))
    fit = (233472 // (smem + 1024)) if smem > 0 else 5
    return 8 if fit >= 8 else (7 if fit == 7 else (6 if fit == 6 else 5))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
