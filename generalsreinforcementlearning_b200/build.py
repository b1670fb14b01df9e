"""Build libgrlcuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["grl_kernels.cu", "grl_abi.cu", "grl_mapgen_gpu.cu", "grl_mapgen.cpp"]
HEADERS = ["grl_layout.h", "grl_launch.h", "grl_mapgen.h", "go_rng_cooked.inc", "go_rng_lehmer_pow.inc", "../../include/grlcuda.h"]
LIB = os.path.join(CSRC, "libgrlcuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-shared",
]


def nvcc_path() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """defines/out build experiment variants (e.g. -DGRL_PERSISTENT=1) next to the product library."""
    if not force and out == LIB and not needs_build():
        return LIB
    cmd = ([nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
           + ["-o", out] + SOURCES)
    proc = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("nvcc failed building libgrlcuda.so")
    if verbose:
        print(proc.stdout)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
