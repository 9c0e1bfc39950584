"""In-tree builds: libflan_b200.so (nvcc, sm_100a) and the CPU thread emulator used by the tests.

    python -m flan_b200.build          # both

nvcc cross-compiles without a GPU; the .so files are git-ignored but travel to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")

NVCC_FLAGS = [
    "-std=c++17", "--expt-relaxed-constexpr",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*names):
    return [os.path.join(CSRC, n) for n in names]


def lib_path():
    return os.path.join(LIB, "libflan_b200.so")


def emu_path():
    return os.path.join(LIB, "libpv_emu.so")


def build_library(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    srcs = _sources("pv_kernels.cu", "pv_capi.cu")
    deps = srcs + _sources("pv_core.cuh", "pv_body.cuh", "pv_tables.h", "pv_launch.h") + \
        [os.path.join(os.path.dirname(HERE), "include", "flan_b200.h")]
    if not force and not _newer(lib_path(), deps):
        return lib_path()
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", lib_path()] + srcs
    subprocess.run(cmd, check=True)
    return lib_path()


def build_emulator(force=False):
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(CSRC, "emu", "pv_emu.cpp")
    deps = [src] + _sources("pv_core.cuh", "pv_body.cuh", "pv_tables.h")
    if not force and not _newer(emu_path(), deps):
        return emu_path()
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-std=c++20", "-fPIC", "-shared", "-I" + cuda_inc,
           "-o", emu_path(), src, "-lpthread"]
    subprocess.run(cmd, check=True)
    return emu_path()


def build_all(force=False):
    build_library(force)
    build_emulator(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(lib_path())
    print(emu_path())
