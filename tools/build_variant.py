"""Builds a variant of the library for A/B timing: one source recompiled with extra nvcc flags
(e.g. -DSPLINE_CLOCKS), linked with the other objects of the regular build.
usage: python tools/build_variant.py NAME SOURCE.cu [nvcc flags...]  ->  sakura_b200/libsakura_b200_NAME.so
Select it at run time with SAKURA_B200_LIB=sakura_b200/libsakura_b200_NAME.so"""
import os, subprocess, sys
sys.path.insert(0, ".")
from sakura_b200 import build as b
b.build()
name, src, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
objdir = os.path.join(b.HERE, "build")
obj = os.path.join(objdir, f"{src}.{name}.o")
cmd = [b._nvcc(), "-ccbin", b._host_cxx()] + b.NVCC_FLAGS + extra + ["-Xptxas", "-v", "-c", os.path.join(b.CSRC, src), "-o", obj]
res = subprocess.run(cmd, capture_output=True, text=True)
sys.stderr.write("\n".join(l for l in res.stderr.splitlines() if "registers" in l or "spill" in l or "error" in l) + "\n")
assert res.returncode == 0, res.stderr
objs = [os.path.join(objdir, s + ".o") for s in b.HOST_SOURCES + b.CUDA_SOURCES if s != src] + [obj]
out = os.path.join(b.HERE, f"libsakura_b200_{name}.so")
subprocess.run([b._nvcc(), "-ccbin", b._host_cxx(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", out] + objs, check=True)
print(out)
