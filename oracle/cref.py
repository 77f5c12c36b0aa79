"""ctypes loader of the oracle's C restatements (CPU baseline / checker; test infrastructure only).

The shared object is rebuilt whenever the hash of (sources, flags, host CPU model) differs from the
stamp written next to it: the library is compiled with -march=native, and a copy built in one
container must not be reused on a box with another CPU (it travels with the gpurun snapshot)."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbruss_ref.so")
SOURCES = ["bruss_ref.c", "configs_ref.c"]
CFLAGS = ["-O3", "-march=native", "-fopenmp", "-fPIC", "-std=c11", "-ffp-contract=off"]


def host_threads():
    """Host cores this process may run on (NOT OMP_NUM_THREADS: torchrun sets that to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _stamp():
    h = hashlib.sha256()
    for s in SOURCES:
        p = os.path.join(HERE, s)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    h.update(" ".join(CFLAGS).encode())
    h.update(_cpu_model().encode())
    return h.hexdigest()


def build(force=False):
    stamp_file = LIB + ".stamp"
    want = _stamp()
    have = open(stamp_file).read().strip() if os.path.exists(stamp_file) and os.path.exists(LIB) else ""
    if force or have != want:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
        tmp = LIB + f".{os.getpid()}.tmp"
        subprocess.check_call(["gcc", *CFLAGS, "-shared", "-o", tmp, *srcs, "-lm"])
        os.replace(tmp, LIB)                      # atomic: several ranks may build at once
        open(stamp_file, "w").write(want)
    return LIB


_lib = None
dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.bruss_ref_rhs.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib.bruss_ref_rhs_slab.argtypes = [dp, dp, dp, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib.bruss_ref_max_threads.restype = C.c_int
        _lib.fisher3d_ref_rhs_slab.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        ip = C.POINTER(C.c_int)
        _lib.burgers2d_ref_rhs.argtypes = [dp, dp, C.c_int, C.c_int] + [dp, ip] * 6 + [dp, dp, dp, dp, C.c_double, C.c_double,
                                                                                  dp, dp, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(dp)


def bruss_rhs(u, xg, yg, N, t, alpha=10.0, nthreads=1, out=None):
    u = np.ascontiguousarray(u, dtype=np.float64)
    du = np.empty_like(u) if out is None else out
    lib().bruss_ref_rhs(_p(du), _p(u), _p(np.ascontiguousarray(xg)), _p(np.ascontiguousarray(yg)),
                        int(N), float(alpha), float(t), int(nthreads))
    return du


def bruss_rhs_slab(u, lo_u, hi_u, lo_v, hi_v, xg, yrows, NX, rows, t, alpha=10.0, nthreads=1):
    """One rank's slab (rows x NX unknowns per species) given the neighbouring slabs' edge rows."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert u.size == 2 * NX * rows and len(yrows) == rows
    du = np.empty_like(u)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (lo_u, hi_u, lo_v, hi_v, xg, yrows)]
    lib().bruss_ref_rhs_slab(_p(du), _p(u), *[_p(a) for a in arrs], int(NX), int(rows), float(alpha), float(t), int(nthreads))
    return du


def fisher3d_rhs_slab(u, lo, hi, NX, NY, planes, h, D=1.0, nthreads=1):
    """Config 5 on one slab of z planes (periodic x, y); lo / hi = the planes below / above the slab."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert u.size == NX * NY * planes
    du = np.empty_like(u)
    lo, hi = np.ascontiguousarray(lo, dtype=np.float64), np.ascontiguousarray(hi, dtype=np.float64)
    lib().fisher3d_ref_rhs_slab(_p(du), _p(u), _p(lo), _p(hi), int(NX), int(NY), int(planes), float(h), float(D), int(nthreads))
    return du


class Burgers2D:
    """Config 3 at benchmark size: the row tables (first tap + weights per node) are taken from the Python oracle's own
    builders for the problem's grids, the evaluation is oracle/configs_ref.c burgers2d_ref_rhs."""

    def __init__(self, pdesys, disc, nu=1.0 / 80):
        from .discretize import OracleProblem
        # a tiny problem on the same x grid / y grid would not give the same per-node rows, so build the oracle's tables
        # for the real grids (cheap: O(n) Fornberg solves), without ever calling its Python RHS
        self.orc = orc = OracleProblem(pdesys, disc)
        self.nx, self.ny = orc.n
        self.nu = nu
        self.tabs = []
        for j, n in enumerate(orc.n):
            dd = orc.dd[j]
            wm, sm = np.zeros((n + 1, 2)), np.ones(n + 1, dtype=np.int32)
            wp, sp_ = np.zeros((n + 1, 2)), np.ones(n + 1, dtype=np.int32)
            w2, s2 = np.zeros((n + 1, 3)), np.ones(n + 1, dtype=np.int32)
            for i in range(2, n):
                w, taps = orc.upwind_row(dd.windneg[1], i, n, True, False, False)        # coef > 0: backward taps
                assert len(w) == 2 and taps[1] == taps[0] + 1
                wm[i], sm[i] = w, taps[0]
                w, taps = orc.upwind_row(dd.windpos[1], i, n, False, False, False)
                assert len(w) == 2 and taps[1] == taps[0] + 1
                wp[i], sp_[i] = w, taps[0]
                w, taps = orc.centered_row(dd.map[2], i, n, False, False)
                w, taps = np.asarray(w, dtype=float), list(taps)
                if len(w) != 3:                      # one-sided rows next to the walls (4 taps at order 2): not on this grid
                    raise NotImplementedError("burgers2d_ref: second-derivative rows with more than 3 taps")
                w2[i], s2[i] = w, taps[0]
            blo, _ = orc.centered_row(dd.map[1], 1, n, False, False)
            bhi, _ = orc.centered_row(dd.map[1], n, n, False, False)
            self.tabs.append((wm, sm, wp, sp_, w2, s2, np.asarray(blo, dtype=float), np.asarray(bhi, dtype=float)))
        self.xg = np.ascontiguousarray(orc.grid[0], dtype=np.float64)
        self.workU = np.zeros(self.nx * self.ny + self.nx)
        self.workV = np.zeros(self.nx * self.ny + self.nx)
        self.nstate = 2 * (self.nx - 2) * (self.ny - 2)

    def rhs(self, u, t, nthreads=1):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.nstate
        du = np.empty_like(u)
        ip = C.POINTER(C.c_int)
        args = []
        for (wm, sm, wp, sp_, w2, s2, blo, bhi) in self.tabs:
            for w, s in ((wm, sm), (wp, sp_), (w2, s2)):
                args += [_p(np.ascontiguousarray(w)), np.ascontiguousarray(s).ctypes.data_as(ip)]
        (_, _, _, _, _, _, bxlo, bxhi), (_, _, _, _, _, _, bylo, _) = self.tabs
        lib().burgers2d_ref_rhs(_p(du), _p(u), self.nx, self.ny, *args, _p(bxlo), _p(bxhi), _p(bylo), _p(self.xg),
                                float(self.nu), float(t), _p(self.workU), _p(self.workV), int(nthreads))
        return du
