"""Build librmgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librmgpu.so")
SOURCES = ["engine.cu", "host_io.cpp", "remote.cpp"]
SERVER = os.path.join(HERE, "rmgpu_server")
DEPS = SOURCES + ["kernels.cuh", "host.hpp", "remote.hpp", "server_main.cpp", "../../include/rmgpu.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-O2,-Wall", "-shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(SERVER):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building librmgpu.so")
    with open(os.path.join(HERE, "librmgpu.ptxas.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    # the persistent worker: a thin main() over the same shared library
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-o", SERVER, os.path.join(CSRC, "server_main.cpp"), "-L" + HERE,
           "-lrmgpu", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building rmgpu_server")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
