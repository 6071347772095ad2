"""Build libvdl2gpu.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

    python -m dumpvdl2_b200.build        # or: from dumpvdl2_b200.build import build_native
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvdl2gpu.so")
SOURCES = ["vdl2_kernels.cu", "vdl2_host.cu", "vdl2_dropin.cu", "vdl2_mg.cu"]
HEADERS = ["vdl2_tables_host.h", "vdl2_core.cuh", "vdl2_fastmath.cuh", "vdl2_kernels.h", "vdl2_types.h", "../../include/vdl2gpu.h", "../../include/vdl2_dropin.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-I", CSRC, "-I", os.path.join(HERE, "..", "include")]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libvdl2gpu.so cannot be built (there is no CPU fallback)")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into dumpvdl2_b200/libvdl2gpu.so; returns its path."""
    if not force and not stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB] + srcs + ["-lpthread", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
