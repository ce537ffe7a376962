"""Build the CUDA library in-tree: simplehydrology_b200/libshx.so (sm_100a only).

nvcc cross-compiles without a GPU.  --fmad=false / -prec-div / -prec-sqrt are part of the
numerical contract (see csrc/shx_math.cuh), not tuning flags.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libshx.so")
SOURCES = [os.path.join(HERE, "csrc", "shx_api.cu"), os.path.join(HERE, "csrc", "shx_multi.cu")]
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cu", ".cuh", ".h"))) + \
       [os.path.join(ROOT, "include", "shx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libshx.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


HOST_EXAMPLE = os.path.join(HERE, "host", "example_erode")
HOST_BENCH = os.path.join(HERE, "host", "bench_bridge")


def _build_host(src_name, out):
    src = os.path.join(HERE, "host", src_name)
    deps = [src, os.path.join(HERE, "host", "shx_world.hpp"), os.path.join(ROOT, "include", "shx.h"), LIB]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", src, "-o", out, "-L" + HERE, "-lshx", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"g++ failed building {src_name}")
    return out


def build_host_example(force=False):
    """the C++ host adaptor's drivers (plain g++, linked with libshx.so through the C ABI only): the example frame
    loop and the end-to-end benchmark bench.py runs"""
    if force:
        for f in (HOST_EXAMPLE, HOST_BENCH):
            if os.path.exists(f):
                os.remove(f)
    _build_host("bench_bridge.cpp", HOST_BENCH)
    return _build_host("example_erode.cpp", HOST_EXAMPLE)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_example(force="--force" in sys.argv))
