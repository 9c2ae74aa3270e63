"""Compile the CUDA engine in-tree for sm_100a (the only target).

    python -m spmv_openmp_cuda_b200.build            # -> spmv_openmp_cuda_b200/lib/libspmv_b200.so

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libspmv_b200.so")
SOURCES = ["engine.cu", "synth.cu"]
DEPS = SOURCES + ["engine_formats.inl", "engine_launch.inl", "engine_multigpu.inl", "engine_hostpath.inl", "engine_shard.inl", "hotx.cuh",
                  "kernels.cuh", "plan.cuh", "xwin.cuh", "common.cuh", "engine.h", os.path.join("..", "..", "include", "spmv_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp,-O2", "--shared", "-lgomp"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DSPMVB200_XW_DEBUG"] if os.environ.get("SPMVB200_XW_DEBUG_BUILD") else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
