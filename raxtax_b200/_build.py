"""In-tree build of the two shared libraries (no JIT cache: the .so files travel with the repo snapshot).

  libraxtax_b200.so  hand-written sm_100a kernels + C ABI (include/raxtax_b200.h)   nvcc
  libraxtax_host.so  C++ host mirror of the reference interface (include/raxtax_host.h)   g++
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
INC = os.path.join(ROOT, "include")
DEVICE_SRC = [os.path.join(PKG, "csrc", "device", f) for f in ("rtx_api.cu", "kernels.cuh", "common.cuh")]
HOST_SRC = [os.path.join(PKG, "csrc", "host", "raxtax_host.cpp")]
DEVICE_LIB = os.path.join(PKG, "libraxtax_b200.so")
HOST_LIB = os.path.join(PKG, "libraxtax_host.so")
CLI_BIN = os.path.join(PKG, "raxtax")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-ldl",
              "-shared"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_device(force=False, verbose=False):
    hdr = [os.path.join(INC, "raxtax_b200.h")]
    if force or _stale(DEVICE_LIB, DEVICE_SRC + hdr):
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", INC, "-o", DEVICE_LIB, DEVICE_SRC[0]]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return DEVICE_LIB


def build_host(force=False):
    hdr = [os.path.join(INC, "raxtax_b200.h"), os.path.join(INC, "raxtax_host.h")]
    if force or _stale(HOST_LIB, HOST_SRC + hdr + [DEVICE_LIB]):
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", INC, "-o", HOST_LIB] + HOST_SRC + [
            "-L", PKG, "-lraxtax_b200", "-lz", "-lpthread", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return HOST_LIB


def build_cli(force=False):
    src = [os.path.join(PKG, "csrc", "host", "main.cpp")]
    if force or _stale(CLI_BIN, src + [HOST_LIB, os.path.join(INC, "raxtax_host.h")]):
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", INC, "-o", CLI_BIN] + src + [
            "-L", PKG, "-lraxtax_host", "-lraxtax_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return CLI_BIN


def build_all(force=False, verbose=False):
    build_device(force, verbose)
    build_host(force)
    build_cli(force)
    return DEVICE_LIB, HOST_LIB


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(DEVICE_LIB)
    print(HOST_LIB)
