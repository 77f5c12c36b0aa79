"""ctypes binding of libmol_cuda.so (include/mol_cuda.h) — the same entry points the Julia
`ccall` shim binds (julia/MOLCuda.jl).  No torch types cross this boundary: raw device
addresses, sizes, host doubles.

The library is REQUIRED: importing this module without a built libmol_cuda.so raises, and every
compute call without a CUDA device returns MOL_E_NOCUDA — there is no CPU fallback anywhere in the
product path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmol_cuda.so")

MOL_OK = 0
MOL_E_NOCUDA = -6
ALG = {"euler": 1, "ssprk33": 2, "rk4": 3, "tsit5": 4}
KERNEL_AUTO, KERNEL_GENERIC = 0, 1
PART_INTERIOR, PART_BOUNDARY = 1, 2


class MolError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmol_cuda error {code}: {msg}")
        self.code = code


class StepStats(C.Structure):
    _fields_ = [("t", C.c_double), ("dt_next", C.c_double), ("eest", C.c_double),
                ("accepted", C.c_int), ("nf", C.c_int)]


class DistInfo(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("halo_planes", C.c_int), ("periodic", C.c_int),
                ("prev_rank", C.c_int), ("next_rank", C.c_int), ("plane_len", C.c_int64), ("first_plane", C.c_int64),
                ("n_planes", C.c_int64), ("state_len_local", C.c_int64), ("state_len_global", C.c_int64),
                ("halo_len", C.c_int64)]


class SolveStats(C.Structure):
    _fields_ = [("t_final", C.c_double), ("dt_last", C.c_double), ("nf", C.c_int64),
                ("naccept", C.c_int64), ("nreject", C.c_int64), ("retcode", C.c_int)]


_lib = None


def _rebuild_if_stale():
    """The library is built in-tree and travels with the source snapshot; if its content stamp does not match
    the sources beside it (edited sources, or a checkout without a binary) and nvcc is present, rebuild it here.
    No nvcc and no library -> lib() raises below: there is nothing to fall back to."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mol_b200_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
        if mod._stale() and os.path.exists(os.path.join(mod.CUDA_HOME, "bin", "nvcc")):
            mod.build_library()
    except RuntimeError:
        raise
    except OSError:
        pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    _rebuild_if_stale()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The CUDA library is the product; there is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
    L.mol_fd_weights.argtypes = [C.c_int, C.c_double, dp, C.c_int, dp]
    L.mol_fd_weights_rows.argtypes = [C.c_int, C.c_int64, C.c_int, dp, dp, dp]
    L.mol_plan_create.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.mol_plan_destroy.argtypes = [vp]
    L.mol_plan_state_len.argtypes = [vp]
    L.mol_plan_state_len.restype = C.c_size_t
    L.mol_plan_nvar.argtypes = [vp]
    L.mol_plan_var_info.argtypes = [vp, C.c_int, C.POINTER(i64), C.POINTER(i64)]
    L.mol_plan_set_option.argtypes = [vp, C.c_char_p, i64]
    L.mol_plan_generated_source.argtypes = [vp]
    L.mol_plan_generated_source.restype = C.c_char_p
    L.mol_plan_cubin.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mol_plan_precompile.argtypes = [vp, C.c_int]
    L.mol_plan_tables.argtypes = [vp, C.POINTER(dp), C.POINTER(C.c_size_t), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_size_t)]
    L.mol_plan_jac_sparsity.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.mol_plan_launch_count.argtypes = [vp]
    L.mol_plan_launch_count.restype = i64
    L.mol_rhs.argtypes = [vp, vp, vp, dp, C.c_double, vp]
    L.mol_rhs_host.argtypes = [vp, vp, vp, dp, C.c_double, C.c_int, vp]
    L.mol_plan_grid_len.argtypes = [vp, C.POINTER(i64)]
    L.mol_plan_grid_len.restype = i64
    L.mol_unpack.argtypes = [vp, vp, vp, C.c_int, dp, dp, vp]
    L.mol_jvp.argtypes = [vp, vp, vp, vp, dp, C.c_double, vp]
    L.mol_rk_init.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
    L.mol_rk_destroy.argtypes = [vp]
    L.mol_rk_set_params.argtypes = [vp, dp]
    L.mol_rk_step.argtypes = [vp, vp, dp, dp, C.c_int, C.POINTER(StepStats), vp]
    L.mol_rk_step_to.argtypes = [vp, vp, vp, dp, dp, C.c_int, C.POINTER(StepStats), vp]
    L.mol_rk_reinit.argtypes = [vp]
    L.mol_rk_solve.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, C.c_int, dp, C.c_int, vp, i64,
                               C.POINTER(SolveStats), vp]
    L.mol_dist_partition.argtypes = [i64, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]
    L.mol_dist_init.argtypes = [vp, C.c_int, C.c_int]
    L.mol_dist_info.argtypes = [vp, C.POINTER(DistInfo)]
    L.mol_dist_unique_id.argtypes = [vp, C.c_size_t]
    L.mol_dist_comm_init.argtypes = [vp, vp, C.c_size_t]
    L.mol_dist_set_halo.argtypes = [vp, vp, vp]
    L.mol_dist_transport.argtypes = [vp]
    L.mol_dist_transport.restype = C.c_char_p
    L.mol_rhs_part.argtypes = [vp, vp, vp, dp, C.c_double, C.c_int, vp]
    L.mol_dist_register.argtypes = [vp, vp]
    L.mol_dist_unregister.argtypes = [vp, vp]
    L.mol_dist_invalidate.argtypes = [vp, vp]
    L.mol_dist_allreduce_sum.argtypes = [vp, vp, C.c_int, vp]
    L.mol_last_error.restype = C.c_char_p
    L.mol_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc):
    if rc != MOL_OK:
        raise MolError(rc, lib().mol_last_error().decode(errors="replace"))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fd_weights(order, x0, x):
    """calculate_weights (fornberg_calculate_weights.jl:20-67) via the library (host code)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.empty(len(x))
    check(lib().mol_fd_weights(int(order), float(x0), _dptr(x), len(x), _dptr(w)))
    return w


def fd_weights_rows(order, x0, x):
    """One row of weights per node: x0 (nrows,), x (nrows, n) -> (nrows, n); each row is fd_weights(order, x0[r], x[r])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    assert x.ndim == 2 and x0.shape == (x.shape[0],)
    w = np.empty_like(x)
    check(lib().mol_fd_weights_rows(int(order), x.shape[0], x.shape[1], _dptr(x0), _dptr(x), _dptr(w)))
    return w


def dist_partition(n_planes, nranks, rank):
    """(first, count) of the planes rank `rank` owns (host arithmetic, no GPU)."""
    a, b = C.c_int64(), C.c_int64()
    check(lib().mol_dist_partition(int(n_planes), int(nranks), int(rank), C.byref(a), C.byref(b)))
    return a.value, b.value


def dist_unique_id():
    """128-byte NCCL unique id (call on rank 0, broadcast to the other ranks)."""
    buf = C.create_string_buffer(128)
    check(lib().mol_dist_unique_id(buf, 128))
    return buf.raw


class Plan:
    """mol_plan handle.  device=-1 compiles only (no GPU needed)."""

    def __init__(self, program: str, device: int = 0):
        self._h = C.c_void_p()
        data = program.encode()
        check(lib().mol_plan_create(data, len(data), int(device), C.byref(self._h)))
        self.device = device
        self.state_len = int(lib().mol_plan_state_len(self._h))
        self.nvar = int(lib().mol_plan_nvar(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mol_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_option(self, key, value):
        check(lib().mol_plan_set_option(self._h, key.encode(), int(value)))

    def generated_source(self):
        return lib().mol_plan_generated_source(self._h).decode()

    def cubin(self, variant):
        p, n = C.c_void_p(), C.c_size_t()
        check(lib().mol_plan_cubin(self._h, variant.encode(), C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def precompile(self, alg: str):
        """Compile every kernel variant of one integrator on several host threads (mol_plan_precompile)."""
        check(lib().mol_plan_precompile(self.handle, {"euler": 1, "ssprk33": 2, "rk4": 3, "tsit5": 4}[alg]))

    def tables(self):
        """(tabw, tabs): copies of the flattened stencil tables as uploaded to the device (introspection)."""
        pw, nw = C.POINTER(C.c_double)(), C.c_size_t()
        ps, ns = C.POINTER(C.c_int)(), C.c_size_t()
        check(lib().mol_plan_tables(self._h, C.byref(pw), C.byref(nw), C.byref(ps), C.byref(ns)))
        tabw = np.ctypeslib.as_array(pw, shape=(nw.value,)).copy() if nw.value else np.zeros(0)
        tabs = np.ctypeslib.as_array(ps, shape=(ns.value,)).copy() if ns.value else np.zeros(0, dtype=np.int32)
        return tabw, tabs

    def jac_sparsity(self):
        """(colptr, rowval) of the Jacobian pattern d(du)/d(u), CSC, 0-based (what `jac_prototype` wants)."""
        nnz = C.c_int64()
        check(lib().mol_plan_jac_sparsity(self._h, None, None, C.byref(nnz)))
        colptr = np.zeros(self.state_len + 1, dtype=np.int64)
        rowval = np.zeros(max(1, nnz.value), dtype=np.int64)
        check(lib().mol_plan_jac_sparsity(self._h, colptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                          rowval.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nnz)))
        return colptr, rowval[:nnz.value]

    def launch_count(self):
        return int(lib().mol_plan_launch_count(self._h))

    def rhs(self, du_ptr, u_ptr, t, p=None, stream=0):
        pp = None if p is None else _dptr(np.ascontiguousarray(p, dtype=np.float64))
        check(lib().mol_rhs(self._h, C.c_void_p(du_ptr), C.c_void_p(u_ptr), pp, float(t), C.c_void_p(stream)))

    def rhs_host(self, du_host_ptr, u_host_ptr, t, p=None, nchunks=0, stream=0):
        """f!(du, u, p, t) on (pinned) HOST buffers: chunked H2D / sweep / D2H pipeline inside the library."""
        pp = None if p is None else _dptr(np.ascontiguousarray(p, dtype=np.float64))
        check(lib().mol_rhs_host(self._h, C.c_void_p(du_host_ptr), C.c_void_p(u_host_ptr), pp, float(t), int(nchunks),
                                 C.c_void_p(stream)))

    def jvp(self, jv_ptr, u_ptr, v_ptr, t, p=None, stream=0):
        """jv = (d f / d u)(u, p, t) v on the device (mol_jvp: forward-mode differentiation of the generated equations)."""
        pp = None if p is None else _dptr(np.ascontiguousarray(p, dtype=np.float64))
        check(lib().mol_jvp(self._h, C.c_void_p(jv_ptr), C.c_void_p(u_ptr), C.c_void_p(v_ptr), pp, float(t), C.c_void_p(stream)))

    def grid_shape(self, ndim):
        """Nodes per dimension of the full grid (boundary nodes included)."""
        n = (C.c_int64 * max(1, ndim))()
        lib().mol_plan_grid_len(self._h, n)
        return tuple(int(v) for v in n[:ndim])

    def unpack(self, full_ptr, u_ptr, times, p=None, stream=0):
        """States -> variables on the whole grid (mol_unpack: sol[u(t,x)] of the reference), all on the device."""
        ts = np.ascontiguousarray(times, dtype=np.float64)
        pp = None if p is None else _dptr(np.ascontiguousarray(p, dtype=np.float64))
        check(lib().mol_unpack(self._h, C.c_void_p(full_ptr), C.c_void_p(u_ptr), len(ts), _dptr(ts), pp, C.c_void_p(stream)))

    # -- slab decomposition (include/mol_cuda.h section e) --------------------------------------------
    def dist_init(self, rank, nranks):
        check(lib().mol_dist_init(self._h, int(rank), int(nranks)))
        self.state_len = int(lib().mol_plan_state_len(self._h))

    def dist_info(self):
        info = DistInfo()
        check(lib().mol_dist_info(self._h, C.byref(info)))
        return info

    def dist_comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(lib().mol_dist_comm_init(self._h, buf, 128))

    def dist_transport(self):
        return lib().mol_dist_transport(self._h).decode()

    def dist_set_halo(self, lo_ptr, hi_ptr):
        check(lib().mol_dist_set_halo(self._h, C.c_void_p(lo_ptr), C.c_void_p(hi_ptr)))

    def rhs_part(self, du_ptr, u_ptr, t, part, p=None, stream=0):
        pp = None if p is None else _dptr(np.ascontiguousarray(p, dtype=np.float64))
        check(lib().mol_rhs_part(self._h, C.c_void_p(du_ptr), C.c_void_p(u_ptr), pp, float(t), int(part), C.c_void_p(stream)))

    def dist_allreduce_sum(self, dev_ptr, n, stream=0):
        check(lib().mol_dist_allreduce_sum(self._h, C.c_void_p(dev_ptr), int(n), C.c_void_p(stream)))


class RK:
    def __init__(self, plan: Plan, alg: str, abstol=1e-6, reltol=1e-3):
        self._h = C.c_void_p()
        self.plan = plan
        check(lib().mol_rk_init(plan.handle, ALG[alg], float(abstol), float(reltol), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mol_rk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, p):
        check(lib().mol_rk_set_params(self._h, _dptr(np.ascontiguousarray(p, dtype=np.float64))))

    def step(self, u_ptr, t, dt, adaptive=False, stream=0):
        tt, dd, st = C.c_double(t), C.c_double(dt), StepStats()
        check(lib().mol_rk_step(self._h, C.c_void_p(u_ptr), C.byref(tt), C.byref(dd), int(adaptive), C.byref(st),
                                C.c_void_p(stream)))
        return tt.value, dd.value, st

    def step_to(self, u_in_ptr, u_out_ptr, t, dt, adaptive=False, stream=0):
        """One step from u_in into the different array u_out (no state copy); returns (t, dt_next, stats)."""
        tt, dd, st = C.c_double(t), C.c_double(dt), StepStats()
        check(lib().mol_rk_step_to(self._h, C.c_void_p(u_in_ptr), C.c_void_p(u_out_ptr), C.byref(tt), C.byref(dd),
                                   int(adaptive), C.byref(st), C.c_void_p(stream)))
        return tt.value, dd.value, st

    def reinit(self):
        check(lib().mol_rk_reinit(self._h))

    def solve(self, u_ptr, t0, t1, dt0=0.0, adaptive=True, saveat=None, save_ptr=0, maxiters=10 ** 6, stream=0):
        st = SolveStats()
        if saveat is None or len(saveat) == 0:
            sp, ns = None, 0
        else:
            sa = np.ascontiguousarray(saveat, dtype=np.float64)
            sp, ns = _dptr(sa), len(sa)
        check(lib().mol_rk_solve(self._h, C.c_void_p(u_ptr), float(t0), float(t1), float(dt0), int(adaptive), sp, ns,
                                 C.c_void_p(save_ptr), int(maxiters), C.byref(st), C.c_void_p(stream)))
        return st
