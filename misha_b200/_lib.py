"""ctypes binding of libmishagpu.so (include/mishagpu.h). Loading fails loudly: there is no CPU fallback."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("MISHAGPU_LIB") or os.path.join(HERE, "libmishagpu.so")  # MISHAGPU_LIB: development builds of the same ABI
HEADER = os.path.join(HERE, "..", "include", "mishagpu.h")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p
c_dbl = ctypes.c_double

_lib = None


class MishaGpuError(RuntimeError):
    """Raised when a libmishagpu call returns non-zero (the host turns it into verror(), src/rdbutils.cpp:720-733)."""


def declared_symbols():
    """Every function name include/mishagpu.h declares (used by the CPU test that checks the exports)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if "MISHAGPU_LIB" not in os.environ:
        from . import build
        build.build()  # no-op when the library is newer than every source (or when there is no nvcc: the prebuilt one is used)
    lib = ctypes.CDLL(SO)
    lib.mg_last_error.restype = ctypes.c_char_p
    lib.mg_last_error.argtypes = [c_vp]
    lib.mg_launch_count.restype = c_i64
    lib.mg_launch_count.argtypes = [c_vp]
    lib.mg_rows_count.restype = c_i64
    lib.mg_rows_count.argtypes = [c_vp]
    lib.mg_dense_bytes.restype = c_i64
    lib.mg_dense_bytes.argtypes = [c_vp]
    for name in ("mg_shutdown",):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [c_vp]
    for name in ("mg_dense_free", "mg_sparse_free", "mg_ivset_free", "mg_seq_free", "mg_rows_free"):
        if hasattr(lib, name):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [c_vp, c_vp]
    lib.mg_group_last_error.restype = ctypes.c_char_p
    lib.mg_group_last_error.argtypes = [c_vp]
    lib.mg_group_ctx.restype = c_vp
    lib.mg_group_ctx.argtypes = [c_vp, c_int]
    lib.mg_group_shutdown.restype = None
    lib.mg_group_shutdown.argtypes = [c_vp]
    lib.mg_group_dense_free.restype = None
    lib.mg_group_dense_free.argtypes = [c_vp, c_vp]
    lib.mg_group_dense_bytes.restype = c_i64
    lib.mg_group_dense_bytes.argtypes = [c_vp, c_int]
    if hasattr(lib, "mg_summary_finalize"):
        lib.mg_summary_finalize.restype = None
    _lib = lib
    return lib
