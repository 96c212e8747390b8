"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE, not product code).

    python -m oracle.build          # builds oracle/_build/liboracle.so

Flags matter for parity: -ffp-contract=off forbids a*b+c fusion so that every
+,-,*,/ is a separately rounded IEEE operation, exactly like the product CUDA
kernels (built with -fmad=false).  -O2 only; no -ffast-math.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "liboracle.so")
SOURCES = ["softras_oracle.c", "nmr_oracle.c"]


def _stale(srcs):
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in srcs + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if not force and not _stale(srcs):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-fno-fast-math", "-Wall", "-Wextra", "-o", LIB] + srcs + ["-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
