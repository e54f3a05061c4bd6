"""Builds libsakura_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m sakura_b200.build [--force]

The shared library exports the C ABI declared in include/sakura_b200.h and nothing
torch-specific; it links the CUDA runtime statically so that it can be dlopen'ed from
any host (C, C++, ctypes, cgo ...).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsakura_b200.so")

CUDA_SOURCES = ["capi_core.cu", "capi_fit.cu", "capi_filters.cu", "capi_stats.cu", "lsqfit_general.cu", "lsqfit_poly.cu", "lsqfit_poly_screen.cu", "lsqfit_poly_warp.cu", "statistics.cu",
                "lsq_blocks.cu", "lsqfit_sinusoid.cu", "lsqfit_spline.cu", "calibration.cu", "bool_filters.cu", "interpolation.cu", "mask_bits.cu"]
HOST_SOURCES = ["basis_tables.cc", "host_bits.cc"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # every fused multiply-add in device code is explicit (see csrc/device_common.cuh)
    "--fmad=false",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsakura_b200.so cannot be built")


def _host_cxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _newest_source_mtime() -> float:
    newest = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in os.listdir(root):
            newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    newest = max(newest, os.path.getmtime(os.path.abspath(__file__)))
    return newest


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the library if it is missing or older than its sources; returns its path."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = _nvcc()
    cxx = _host_cxx()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    def compile_cuda(src: str) -> str:
        obj = os.path.join(objdir, src + ".o")
        # a source whose object is newer than every source file (headers included) is not recompiled
        newest = max(headers_mtime, os.path.getmtime(os.path.join(CSRC, src)))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            return obj
        cmd = [nvcc, "-ccbin", cxx] + NVCC_FLAGS + ["-Xptxas", "-v", "-c",
                                                    os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, check=False, capture_output=True, text=True)
        with open(os.path.join(objdir, src + ".ptxas.log"), "w") as fh:
            fh.write(res.stderr)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed on " + src)
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    headers_mtime = max([os.path.getmtime(os.path.abspath(__file__))]
                        + [os.path.getmtime(os.path.join(root, name))
                           for root in (CSRC, os.path.join(HERE, "..", "include")) for name in os.listdir(root)
                           if name.endswith((".h", ".cuh"))])
    with ThreadPoolExecutor(max_workers=min(len(CUDA_SOURCES), os.cpu_count() or 4)) as pool:
        objs += list(pool.map(compile_cuda, CUDA_SOURCES))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-ccbin", cxx, "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
           "-cudart", "static", "-o", tmp] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(path)
