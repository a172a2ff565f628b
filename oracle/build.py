"""Build the oracle's C restatement (TEST INFRASTRUCTURE) into oracle/_build/libo2a_oracle.so.

The reference is pure Python: there is no C/C++ reference source to compile, so
there is no `oracle/_ref/` binary; `oracle/ref_harness.py` imports the Python reference in the
build container instead, and `tests/golden/` carries its outputs to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libo2a_oracle.so")
SRC = os.path.join(HERE, "o2a_oracle.c")
SRCS = [SRC, os.path.join(HERE, "o2a_oracle_producer.c"), os.path.join(HERE, "o2a_oracle_annotate.c")]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in SRCS):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-Wno-unused-function", "-o", LIB] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
