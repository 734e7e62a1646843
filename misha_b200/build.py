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
    "--shared", "-Xcompiler", "-fPIC,-O3",
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


HOST = os.path.join(HERE, "host")
HOST_SO = os.path.join(HERE, "libmishahost.so")
HOST_DEMO = os.path.join(HERE, "mg_host_demo")
CXX = os.environ.get("CXX", "g++")


def build_host(force=False):
    """The C++ host layer above the C ABI (misha_b200/host): libmishahost.so (GpuTrackExprScanner, the mirror of the
    reference's TrackExprScanner for this path) and the mg_host_demo driver the tests run."""
    srcs = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [SO]
    if not force and all(os.path.exists(p) and all(os.path.getmtime(s) <= os.path.getmtime(p) for s in srcs) for p in (HOST_SO, HOST_DEMO)):
        return HOST_SO
    import shutil
    if shutil.which(CXX) is None:
        if os.path.exists(HOST_SO) and os.path.exists(HOST_DEMO):
            return HOST_SO  # GPU box without a compiler: the prebuilt files travelled with the snapshot
        raise RuntimeError("%s not found and no prebuilt host layer" % CXX)
    common = [CXX, "-std=c++17", "-O2", "-Wall"]
    link = ["-L" + HERE, "-lmishagpu", "-Wl,-rpath,$ORIGIN"]
    with _BuildLock():
        tmp_so, tmp_demo = HOST_SO + ".tmp.%d" % os.getpid(), HOST_DEMO + ".tmp.%d" % os.getpid()
        try:
            subprocess.check_call(common + ["-fPIC", "-shared", "-o", tmp_so, os.path.join(HOST, "GpuTrackExprScanner.cpp")] + link)
            os.replace(tmp_so, HOST_SO)
            subprocess.check_call(common + ["-o", tmp_demo, os.path.join(HOST, "mg_host_demo.cpp"), "-L" + HERE, "-lmishahost"] + link[1:])
            os.replace(tmp_demo, HOST_DEMO)
        finally:
            for t in (tmp_so, tmp_demo):
                if os.path.exists(t):
                    os.remove(t)
    return HOST_SO


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    return all(os.path.getmtime(d) <= t for d in deps())


class _BuildLock:
    """One builder at a time per checkout: N ranks of a torchrun job import the package at the same moment, and a library that is being
    written must never be the one another process loads. The lock file lives next to the library; whoever gets it re-checks the
    timestamps (the previous holder may have built already)."""

    def __enter__(self):
        import fcntl
        self.f = open(os.path.join(HERE, ".build.lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    if not os.path.exists(NVCC):
        if os.path.exists(SO):
            return SO  # GPU box without a toolkit: use the prebuilt library that travelled with the snapshot
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, SO))
    with _BuildLock():
        if not force and up_to_date():
            return SO
        tmp = SO + ".tmp.%d" % os.getpid()  # written aside, renamed into place: a reader sees the old library or the new one, never half of one
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
        if verbose:
            print(" ".join(cmd))
        try:
            subprocess.check_call(cmd)
            os.replace(tmp, SO)
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    build_host(force="--force" in sys.argv)
    print(SO)
