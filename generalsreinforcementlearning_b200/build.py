"""Build libgrlcuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every source is compiled to an object in parallel (one nvcc process each), then linked."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["grl_turn_20.cu", "grl_turn_15.cu", "grl_turn_10.cu", "grl_turn_generic.cu", "grl_kernels.cu", "grl_abi.cu",
           "grl_mapgen_gpu.cu", "grl_mapgen.cpp", "grl_expand.cpp"]
HEADERS = ["grl_layout.h", "grl_launch.h", "grl_mapgen.h", "grl_device.cuh", "grl_obs.cuh", "grl_gym.cuh", "grl_turn.cuh", "go_rng_cooked.inc", "go_rng_lehmer_pow.inc", "../../include/grlcuda.h"]
LIB = os.path.join(CSRC, "libgrlcuda.so")
OBJ_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
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


def source_hash() -> str:
    """sha256 over the sources and headers the library is built from (bench.py keys profiles/traffic.json by it)."""
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


TURN_UNIT = ["grl_turn_20.cu", "grl_launch.h", "grl_turn.cuh", "grl_device.cuh", "grl_gym.cuh", "grl_obs.cuh", "grl_layout.h",
             "../../include/grlcuda.h"]


def turn_source_hash() -> str:
    """sha256 over the translation unit of the headline turn kernel (grl_turn_20.cu and every header it includes):
    profiles/traffic.json is keyed by it, so a change elsewhere in the library (ABI glue, off-path kernels) does not
    void the capture while any change the kernel could see does."""
    h = hashlib.sha256()
    for f in sorted(TURN_UNIT):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def _run(cmd, verbose):
    proc = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))
    if verbose:
        print(proc.stdout)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """defines/out build experiment variants next to the product library."""
    if not force and out == LIB and not needs_build():
        return LIB
    tag = hashlib.sha256(" ".join(defines).encode()).hexdigest()[:8] if defines else "product"
    odir = os.path.join(OBJ_DIR, tag)
    os.makedirs(odir, exist_ok=True)
    nvcc = nvcc_path()
    flags = NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    jobs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(odir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(os.path.join(CSRC, src)))
        if stale:
            jobs.append([nvcc] + flags + ["-c", "-o", obj, src])
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(jobs))) as pool:
        list(pool.map(lambda c: _run(c, verbose), jobs))
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, verbose)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
