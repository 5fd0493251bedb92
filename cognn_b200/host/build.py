"""Builds libcognn_b200_host.so: the C++ host engine above the C ABI (links libcognn_b200.so, cudart, NCCL)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libcognn_b200_host.so")
SOURCES = ["engine.cpp", "comm.cpp", "capi_engine.cpp"]


def _cxx():
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def build(force=False):
    deps = [os.path.join(HERE, s) for s in SOURCES] + [os.path.join(HERE, "engine.h"),
                                                       os.path.join(PKG, "..", "include", "cognn_b200_engine.h"),
                                                       os.path.join(PKG, "..", "include", "cognn_b200.h"),
                                                       os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = [_cxx(), "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-fopenmp", "-o", LIB] + \
          [os.path.join(HERE, s) for s in SOURCES] + \
          ["-I/usr/local/cuda/include", "-L" + PKG, "-l:libcognn_b200.so", "-Wl,-rpath,$ORIGIN",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl"]
    subprocess.check_call(cmd)
    return LIB


DEMO = os.path.join(PKG, "shim_demo")


def build_shim_demo(force=False):
    """tests/shim_demo.cpp against the header-only shim (cognn_b200/host/shim/cognn_shim.h) and libcognn_b200.so."""
    src = os.path.join(PKG, "..", "tests", "shim_demo.cpp")
    deps = [src, os.path.join(HERE, "shim", "cognn_shim.h"), os.path.join(PKG, "..", "include", "cognn_b200.h")]
    if not force and os.path.exists(DEMO) and all(os.path.getmtime(DEMO) >= os.path.getmtime(d) for d in deps):
        return DEMO
    cmd = [_cxx(), "-O2", "-g", "-std=c++17", "-Wall", "-pthread", "-o", DEMO, src, "-L" + PKG, "-l:libcognn_b200.so",
           "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)
    return DEMO


HARNESS = os.path.join(PKG, "gcn-optimize-b200")


def build_harness(force=False):
    """cognn_b200/host/harness.cpp: the reference's CLI (harness.cpp / harness.h) on top of the engine."""
    src = os.path.join(HERE, "harness.cpp")
    deps = [src, os.path.join(HERE, "engine.h"), LIB]
    if not force and os.path.exists(HARNESS) and all(os.path.getmtime(HARNESS) >= os.path.getmtime(d) for d in deps):
        return HARNESS
    cmd = [_cxx(), "-O2", "-g", "-std=c++17", "-Wall", "-Wno-implicit-fallthrough", "-pthread", "-o", HARNESS, src, "-L" + PKG,
           "-l:libcognn_b200_host.so", "-l:libcognn_b200.so", "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl",
           "-fopenmp"]
    subprocess.check_call(cmd)
    return HARNESS


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_shim_demo(force="--force" in sys.argv))
    print(build_harness(force="--force" in sys.argv))
