"""Build libo2a_cuda.so (sm_100a) in-tree with nvcc: `python -m ortho2align_b200.build [--force]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libo2a_cuda.so")
SOURCES = [os.path.join(CSRC, "o2a_api.cu")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false",                       # never contract a*b+c: bits are compared with numpy/scipy
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


INGEST_LIB = os.path.join(LIB_DIR, "libo2a_ingest.so")
INGEST_SRC = os.path.join(CSRC, "o2a_ingest.cpp")


def build_ingest(force=False):
    """Host-only C++ JSON ingest library (g++)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    inc = os.path.join(os.path.dirname(HERE), "include")
    # JSON ingest + columnar packer + BLAST table reader + record formatter, one library
    srcs = [INGEST_SRC, os.path.join(CSRC, "o2a_pack.cpp"), os.path.join(CSRC, "o2a_blast.cpp"), os.path.join(CSRC, "o2a_records.cpp")]
    deps = srcs + [os.path.join(CSRC, "o2a_pack.h"), os.path.join(inc, "o2a_ingest.h"), os.path.join(inc, "o2a_blast.h"), os.path.join(inc, "o2a_records.h")]
    if not force and os.path.exists(INGEST_LIB) and all(os.path.getmtime(INGEST_LIB) >= os.path.getmtime(d) for d in deps):
        return INGEST_LIB
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", INGEST_LIB] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libo2a_ingest.so")
    return INGEST_LIB


def _deps():
    out = list(SOURCES)
    for fn in os.listdir(CSRC):
        if fn.endswith((".cuh", ".h")):
            out.append(os.path.join(CSRC, fn))
    out.append(os.path.join(os.path.dirname(HERE), "include", "o2a.h"))
    return out


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libo2a_cuda.so")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_ingest(force="--force" in sys.argv))
