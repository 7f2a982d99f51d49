"""Builds the CUDA extension in-tree: mapad_b200/libmapad_gpu.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmapad_gpu.so")
SOURCES = ["mapad_gpu.cu", "gpu_index_build.cu", "host_index.cpp", "host_params.cpp", "dev_index_build.cpp", "host_io.cpp", "host_index_files.cpp"]
HEADERS = ["common.h", "dev_index.cuh", "search_core.cuh", "epilogue_core.cuh", "libm_emu.cuh", "host_index.hpp",
           "host_params.hpp", "dev_index_build.hpp", "sais.hpp", "search_group.cuh", "simt.cuh", "../../include/mapad_gpu.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # exact f32 parity with the reference: never contract a*b+c (every FMA in the source is explicit)
    "--fmad=false", "--Werror", "cross-execution-space-call", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "static",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: build a tuning variant (extra nvcc flags, e.g. -DMAPAD_POOL_MIN_BLOCKS=6) next to the product library;
    a variant is selected at run time with MAPAD_GPU_LIB=<path>."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = ([nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] +
           [os.path.join(CSRC, f) for f in SOURCES] + ["-lz"])
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
