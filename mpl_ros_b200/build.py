"""In-tree build of libmplb.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "mplb.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "mplb_search.cuh"), os.path.join(HERE, "csrc", "mplb_device.cuh"),
        os.path.join(HERE, "csrc", "mplb_trig.cuh"),
        os.path.join(HERE, "..", "include", "mplb.h")]
OUT = os.path.join(HERE, "libmplb.so")

# -fmad=false: the reference is built without FMA contraction (MPL/CMakeLists.txt:5-8); the kernels also use
# explicit __d*_rn intrinsics, the flag covers whatever remains.
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-fmad=false"]


def build_lib(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT
