"""Builds misha_b200/libmishagpu.so (the C-ABI library of include/mishagpu.h) in-tree for sm_100a.

    python -m misha_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmishagpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-O2",
    "-cudart", "static",
    # IEEE arithmetic everywhere: parity with the reference's libm/FPU results is the first gate
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
]


def sources():
    return [os.path.join(CSRC, "mishagpu.cu")]


def deps():
    out = [os.path.join(HERE, "..", "include", "mishagpu.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    return all(os.path.getmtime(d) <= t for d in deps())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    if not os.path.exists(NVCC):
        if os.path.exists(SO):
            return SO  # GPU box without a toolkit: use the prebuilt library that travelled with the snapshot
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, SO))
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + sources()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(SO)
