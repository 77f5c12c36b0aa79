"""ctypes loader of oracle/bruss_ref.c (CPU baseline; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbruss_ref.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "bruss_ref.c")):
        subprocess.check_call(["make", "-C", HERE, "-B" if force else "-s", "_build/libbruss_ref.so"],
                              stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        _lib.bruss_ref_rhs.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib.bruss_ref_max_threads.restype = C.c_int
    return _lib


def bruss_rhs(u, xg, yg, N, t, alpha=10.0, nthreads=1, out=None):
    u = np.ascontiguousarray(u, dtype=np.float64)
    du = np.empty_like(u) if out is None else out
    dp = C.POINTER(C.c_double)
    lib().bruss_ref_rhs(du.ctypes.data_as(dp), u.ctypes.data_as(dp),
                        np.ascontiguousarray(xg).ctypes.data_as(dp), np.ascontiguousarray(yg).ctypes.data_as(dp),
                        int(N), float(alpha), float(t), int(nthreads))
    return du
