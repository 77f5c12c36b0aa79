"""Runs the generated table-driven kernel of a stencil program on the CPU (g++ + tests/cuda_emu/cuda_emu.h)."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

WRAPPER = r'''
// ---- emulation entry points (appended by tests/cuda_emu) ------------------------------------------------------------
// slab decomposition (MOL_DIST = 1): this rank's planes of the split dimension, doubles per variable, ghost planes per input
static int emu_loc_lo = 0, emu_loc_hi = -1;
static long long emu_vstride = 0;
static const double* emu_hlo[8];
static const double* emu_hhi[8];
extern "C" void emu_set_slab(int lo, int hi, long long vstride, const double* const* hlo, const double* const* hhi, int nin) {
    emu_loc_lo = lo; emu_loc_hi = hi; emu_vstride = vstride;
    for (int j = 0; j < nin; ++j) { emu_hlo[j] = hlo[j]; emu_hhi[j] = hhi[j]; }
}
static const double* emu_ctl = nullptr;       // MOL_DEVDT: the device-side step control block {t, dt, skip}
extern "C" void emu_set_ctl(const double* ctl) { emu_ctl = ctl; }
static void emu_in(MolIn& in, const double* const* arrs, const double* coefs) {
#if MOL_DEVDT
    in.ctl = emu_ctl;
#endif
    for (int j = 0; j < MOL_NIN; ++j) {
        in.a[j] = arrs[j];
        in.c[j] = coefs[j];
#if MOL_DIST
        in.hlo[j] = emu_hlo[j];
        in.hhi[j] = emu_hhi[j];
#endif
    }
}
static void emu_ctx(MolCtx& c, double t, const double* p, const double* const* grid, const double* tabw, const int* tabs) {
    c.t = t;
    for (int k = 0; k < (MOL_NPARAM > 0 ? MOL_NPARAM : 1); ++k) c.p[k] = (MOL_NPARAM > 0) ? p[k] : 0.0;
    for (int j = 0; j < 3; ++j) c.grid[j] = grid[j];
    c.tabw = tabw;
    c.tabs = tabs;
    c.loc_lo = MOL_ILO(0, MOL_NDIM - 1);
    c.loc_hi = MOL_IHI(0, MOL_NDIM - 1);
    c.vstride = 0;
#if MOL_DIST
    c.loc_lo = emu_loc_lo;
    c.loc_hi = emu_loc_hi;
    c.vstride = emu_vstride;
#endif
}
// several CTAs, one after the other (the ticket queue and the static FIN assignment both work sequentially): grid size
static int emu_grid = 1;
extern "C" void emu_set_grid(int g) { emu_grid = g > 0 ? g : 1; }
template <class F>
static void emu_launch_grid(F&& kernel) {
    for (int b = 0; b < emu_grid; ++b) {
        blockIdx.x = b;
        gridDim.x = emu_grid;
        emu_launch(kernel);
    }
    blockIdx.x = 0;
    gridDim.x = 1;
}
static const double* emu_jv = nullptr;
extern "C" void emu_set_jv(const double* v) { emu_jv = v; }
#if MOL_KERNEL_TILED
unsigned char mol_smem_raw[256 * 1024];
// the tiled kernel (cooperative-loader staging) on one box of nodes, one CTA drawing every tile from the ticket queue
extern "C" void emu_rhs(const double* const* arrs, const double* coefs, double t, const double* p, const double* const* grid,
                        const double* tabw, const int* tabs, const int* box, double* out, void* epi_args) {
    MolIn in;
    emu_in(in, arrs, coefs);
    MolCtx c;
    emu_ctx(c, t, p, grid, tabw, tabs);
    MolTiles T;
    memset(&T, 0, sizeof T);
    const int tdim[3] = {MOL_TX, MOL_TY, MOL_TZ};
    int nt[3] = {1, 1, 1}, n = 1;
    for (int j = 0; j < MOL_NDIM; ++j) { nt[j] = (box[3 + j] - box[j] + 1 + tdim[j] - 1) / tdim[j]; n *= nt[j]; }
    T.b[0].nt0 = nt[0]; T.b[0].nt1 = nt[1]; T.b[0].nt2 = nt[2]; T.b[0].ntiles = n;
    for (int j = 0; j < 3; ++j) { T.b[0].lo[j] = box[j]; T.b[0].hi[j] = box[3 + j]; T.b[1].lo[j] = 1; }
    T.b[1].nt0 = T.b[1].nt1 = T.b[1].nt2 = 1;
    T.ntiles = n;
    static int counter;
    counter = 0;
    T.counter = &counter;
#if MOL_TMA
    // stand-in tensor maps (kernels/mol_tiled.cuh, MOL_HOST_EMU): the stored box of every variable, tile-sized boxes
    MolTileMaps maps;
    memset(&maps, 0, sizeof maps);
    for (int v = 0; v < MOL_NVAR; ++v) {
        MolEmuMap* m = reinterpret_cast<MolEmuMap*>(maps.m[v].bytes);
        m->base = arrs[0] + MOL_VOFF(v, c);
        const long long e[3] = {mol_ext_[v][0], mol_ext_[v][1], mol_ext_[v][2]};
        for (int j = 0; j < 3; ++j) m->dim[j] = (j < MOL_NDIM) ? e[j] : 1;
#if MOL_DIST
        m->dim[MOL_NDIM - 1] = c.loc_hi - c.loc_lo + 1;
#endif
        m->stride[0] = 1; m->stride[1] = m->dim[0]; m->stride[2] = m->dim[0] * m->dim[1];
        m->box[0] = MOL_SX; m->box[1] = MOL_SY; m->box[2] = MOL_SZ;
    }
#if MOL_EPI
    const MolEpi epi = *reinterpret_cast<MolEpi*>(epi_args);
    emu_launch_grid([&]() { mol_rhs_tiled(in, c, T, out, maps, epi); });
#else
    emu_launch_grid([&]() { mol_rhs_tiled(in, c, T, out, maps); });
#endif
#else
#if MOL_EPI
    const MolEpi epi = *reinterpret_cast<MolEpi*>(epi_args);
    emu_launch_grid([&]() { mol_rhs_tiled(in, c, T, out, epi); });
#elif MOL_KERNEL_JVP
    MolJv jv;                        // tiled J*v: out = (df/du)(u) v, v set through emu_set_jv
    jv.v = emu_jv;
    emu_launch_grid([&]() { mol_rhs_tiled(in, c, T, out, jv); });
#else
    emu_launch_grid([&]() { mol_rhs_tiled(in, c, T, out); });
#endif
#endif
}
extern "C" int emu_epi_size() {
#if MOL_EPI
    return (int)sizeof(MolEpi);
#else
    return 0;
#endif
}
#elif !MOL_KERNEL_UNPACK && !MOL_KERNEL_SOLVE
extern "C" void emu_rhs(const double* const* arrs, const double* coefs, double t, const double* p, const double* const* grid,
                        const double* tabw, const int* tabs, const int* box, double* out, void* epi_args) {
    MolIn in;
    emu_in(in, arrs, coefs);
    MolCtx c;
    emu_ctx(c, t, p, grid, tabw, tabs);
    MolBoxes B;
    memset(&B, 0, sizeof B);
    B.n = 1;
    mol_i64 total = 1;
    for (int j = 0; j < 3; ++j) { B.b[0].lo[j] = box[j]; B.b[0].hi[j] = box[3 + j]; total *= (box[3 + j] - box[j] + 1); }
    B.start[0] = 0;
    for (int k = 1; k <= MOL_MAX_BOXES; ++k) B.start[k] = total;
#if MOL_EPI
    const MolEpi epi = *reinterpret_cast<MolEpi*>(epi_args);
    emu_launch_grid([&]() { mol_rhs_generic(in, c, B, out, epi); });
#else
    emu_launch_grid([&]() { mol_rhs_generic(in, c, B, out); });
#endif
}
extern "C" int emu_epi_size() {
#if MOL_EPI
    return (int)sizeof(MolEpi);
#else
    return 0;
#endif
}
#elif 0
#endif
#if MOL_KERNEL_JVP && !MOL_KERNEL_TILED
extern "C" void emu_jvp(const double* u, const double* v, double t, const double* p, const double* const* grid, const double* tabw,
                        const int* tabs, const int* box, double* out) {
    MolIn in;
    in.a[0] = u;
    in.c[0] = 1.0;
    MolJv jv;
    jv.v = v;
    MolCtx c;
    emu_ctx(c, t, p, grid, tabw, tabs);
    MolBoxes B;
    memset(&B, 0, sizeof B);
    B.n = 1;
    mol_i64 total = 1;
    for (int j = 0; j < 3; ++j) { B.b[0].lo[j] = box[j]; B.b[0].hi[j] = box[3 + j]; total *= (box[3 + j] - box[j] + 1); }
    for (int k = 1; k <= MOL_MAX_BOXES; ++k) B.start[k] = total;
    emu_launch([&]() { mol_jvp_generic(in, jv, c, B, out); });
}
#endif
#if MOL_KERNEL_SOLVE
// the persistent solver kernel (kernels/mol_generic.cuh): one emulated CTA runs a whole solve
extern "C" void emu_solve(double* u, double* work /* 9 x n */, double* save, const double* saveat, int nsave, int alg, int adaptive,
                          double t0, double t1, double dt0, double abstol, double reltol, long long maxiters, long long n,
                          const double* p, const double* const* grid, const double* tabw, const int* tabs, const int* box,
                          double* out /* 7 */) {
    MolCtx c;
    emu_ctx(c, t0, p, grid, tabw, tabs);
    MolBoxes B;
    memset(&B, 0, sizeof B);
    B.n = 1;
    mol_i64 total = 1;
    for (int j = 0; j < 3; ++j) { B.b[0].lo[j] = box[j]; B.b[0].hi[j] = box[3 + j]; total *= (box[3 + j] - box[j] + 1); }
    for (int k = 1; k <= MOL_MAX_BOXES; ++k) B.start[k] = total;
    MolSolveArgs A;
    memset(&A, 0, sizeof A);
    A.u = u;
    for (int k = 0; k < 9; ++k) A.w[k] = work + (long long)k * n;
    A.save = save; A.saveat = saveat; A.out = out;
    A.t0 = t0; A.t1 = t1; A.dt0 = dt0; A.abstol = abstol; A.reltol = reltol;
    A.maxiters = maxiters; A.n = n; A.nglobal = n;
    A.nsave = nsave; A.alg = alg; A.adaptive = adaptive;
    emu_launch([&]() { mol_solve_small(c, B, A); });
}
#endif
#if MOL_KERNEL_UNPACK
extern "C" void emu_unpack(const double* u, double t, const double* p, const double* const* grid, const double* tabw,
                           const int* tabs, double* out) {
    MolIn in;
    in.a[0] = u;
    in.c[0] = 1.0;
    MolCtx c;
    emu_ctx(c, t, p, grid, tabw, tabs);
    emu_launch([&]() { mol_unpack_full(in, c, out); });
}
#endif
'''


def _shim_header(cache_dir, nthreads):
    """cuda_emu.h, precompiled once per EMU_THREADS value (its <thread> / <barrier> / <functional> includes are half of
    every variant's compile time, and a cold suite compiles ~250 variants): a copy of the header next to its .gch in the
    cache directory, keyed by the header's hash; g++ picks the .gch up through `-include <copy>`."""
    src = open(os.path.join(HERE, "cuda_emu.h")).read()
    tag = hashlib.sha1((src + str(nthreads)).encode()).hexdigest()[:12]
    pdir = os.path.join(cache_dir, f"pch_{tag}")
    hdr = os.path.join(pdir, "cuda_emu.h")
    if not os.path.exists(hdr + ".gch"):
        os.makedirs(pdir, exist_ok=True)
        open(hdr, "w").write(src)
        tmp = hdr + f".{os.getpid()}.gch.tmp"
        r = subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-w", "-x", "c++-header", f"-DEMU_THREADS={nthreads}", hdr,
                            "-o", tmp], capture_output=True, text=True)
        if r.returncode == 0:
            os.replace(tmp, hdr + ".gch")
        else:                                   # no PCH: the plain header still works
            return os.path.join(HERE, "cuda_emu.h")
    return hdr


class EmuKernel:
    def __init__(self, plan, prog, nin=1, epi=0, unpack=False, tiled=False, halo=0, staging="coop", jvp=False, extra_defs=(),
                 solve=False):
        """tiled=True: the tiled kernel, 256 emulated threads, on the core box."""
        gen = plan.generated_source()
        src = gen.replace("extern __shared__ __align__(128) unsigned char mol_smem_raw[];",
                          "extern unsigned char mol_smem_raw[];") + WRAPPER
        nthreads = 32
        if solve:
            nin, nthreads = 7, 64
            extra_defs = list(extra_defs) + ["MOL_KERNEL_SOLVE=1"]
        if tiled:
            import re
            nthreads = int(re.search(r"#define MOL_NTHREADS (\d+)", gen).group(1))
        # staging: "coop" = cooperative loader; "tma" / "cpasync" = the multi-stage pipelines with synchronous host
        # stand-ins for the copy instructions (kernels/mol_tiled.cuh, MOL_HOST_EMU)
        assert staging in ("coop", "tma", "cpasync") and (staging == "coop" or (tiled and nin == 1))
        defs = [f"-DMOL_NIN={nin}", f"-DMOL_EPI={epi}", f"-DMOL_KERNEL_TILED={1 if tiled else 0}",
                f"-DMOL_TMA={1 if staging == 'tma' else 0}", f"-DMOL_CPASYNC={1 if staging == 'cpasync' else 0}", "-DMOL_HOST_EMU=1",
                f"-DMOL_KERNEL_UNPACK={1 if unpack else 0}", f"-DMOL_KERNEL_JVP={1 if jvp else 0}", f"-DEMU_THREADS={nthreads}",
                "-DMOL_MIN_CTAS=1"] + [f"-D{d}" for d in extra_defs]
        if halo:
            defs += ["-DMOL_DIST=1", f"-DMOL_HALO={halo}"]
        key = hashlib.sha1((src + " ".join(defs)).encode()).hexdigest()[:16]
        d = os.path.join(tempfile.gettempdir(), "mol_cuda_emu")
        os.makedirs(d, exist_ok=True)
        so = os.path.join(d, key + ".so")
        if not os.path.exists(so):              # (several test workers may want the same variant: private names, atomic rename)
            cu = os.path.join(d, f"{key}.{os.getpid()}.cpp")
            tmp = os.path.join(d, f"{key}.{os.getpid()}.so.tmp")
            open(cu, "w").write(src)
            cmd = ["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-w", "-include", _shim_header(d, nthreads), *defs,
                   cu, "-o", tmp]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed on the generated source:\n" + r.stderr[-4000:])
            os.replace(tmp, so)
            os.remove(cu)
        self.lib = C.CDLL(so)
        self.prog, self.plan, self.nin, self.epi, self.tiled = prog, plan, nin, epi, tiled
        self.tabw, self.tabs = plan.tables()
        self.tabs = np.ascontiguousarray(self.tabs, dtype=np.int32)
        self.grids = [np.ascontiguousarray(ax.x, dtype=np.float64) for ax in prog.axes]
        while len(self.grids) < 3:
            self.grids.append(np.zeros(1))
        if tiled:
            lo, hi = list(prog.corebox[0]), list(prog.corebox[1])
        else:
            lo = [min(prog.ilo[v][j] for v in range(len(prog.ilo))) for j in range(len(prog.axes))]
            hi = [max(prog.ihi[v][j] for v in range(len(prog.ihi))) for j in range(len(prog.axes))]
        self.box = np.array(lo + [1] * (3 - len(lo)) + hi + [1] * (3 - len(hi)), dtype=np.int32)

    def _common(self, t, p):
        dp = C.POINTER(C.c_double)
        p = np.ascontiguousarray(self.prog.pvals if p is None else p, dtype=np.float64)
        if p.size == 0:
            p = np.zeros(1)
        garr = (dp * 3)(*[g.ctypes.data_as(dp) for g in self.grids])
        return dp, p, garr

    def set_slab(self, loc_lo, loc_hi, vstride, halos_lo, halos_hi):
        """Slab mode (kernels compiled with halo > 0): this rank's plane range of the split dimension, doubles per
        variable, and the ghost planes (below / above the slab) of every input array."""
        dp = C.POINTER(C.c_double)
        self._hl = [np.ascontiguousarray(h, dtype=np.float64) for h in halos_lo]
        self._hh = [np.ascontiguousarray(h, dtype=np.float64) for h in halos_hi]
        lo = (dp * len(self._hl))(*[h.ctypes.data_as(dp) for h in self._hl])
        hi = (dp * len(self._hh))(*[h.ctypes.data_as(dp) for h in self._hh])
        self.lib.emu_set_slab(int(loc_lo), int(loc_hi), C.c_longlong(int(vstride)), lo, hi, len(self._hl))

    def set_ctl(self, t, dt, skip=0.0):
        """MOL_DEVDT variants (extra_defs=["MOL_DEVDT=1"]): the device-side step control block {t, dt, skip}."""
        self._ctl = np.array([t, dt, skip], dtype=np.float64)
        self.lib.emu_set_ctl(self._ctl.ctypes.data_as(C.POINTER(C.c_double)))

    def rhs(self, arrays, coefs, t, p=None, epi_struct=None, nout=None, box=None, grid=1):
        """grid: number of CTAs, run one after the other (tile queue / static FIN assignment, one error slot per CTA)."""
        self.lib.emu_set_grid(int(grid))
        dp, p, garr = self._common(t, p)
        if box is not None:
            nd = len(self.prog.axes)
            self.box = np.array(list(box[:nd]) + [1] * (3 - nd) + list(box[nd:]) + [1] * (3 - nd), dtype=np.int32)
        arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
        aarr = (dp * len(arrays))(*[a.ctypes.data_as(dp) for a in arrays])
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        out = np.zeros(arrays[0].size if nout is None else nout)
        self.lib.emu_rhs(aarr, coefs.ctypes.data_as(dp), C.c_double(t), p.ctypes.data_as(dp), garr,
                         self.tabw.ctypes.data_as(dp) if self.tabw.size else None,
                         self.tabs.ctypes.data_as(C.POINTER(C.c_int)) if self.tabs.size else None,
                         self.box.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(dp),
                         None if epi_struct is None else C.byref(epi_struct))
        self._keep = (arrays, coefs, p, garr, aarr)
        return out

    def jvp(self, u, v, t, p=None):
        if self.tiled:               # tiled J*v: the tiled kernel compiled on dual numbers, on the core box
            dpd = C.POINTER(C.c_double)
            self._jv = np.ascontiguousarray(v, dtype=np.float64)
            self.lib.emu_set_jv(self._jv.ctypes.data_as(dpd))
            return self.rhs([u], [1.0], t, p=p)
        dp, p, garr = self._common(t, p)
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros(u.size)
        self.lib.emu_jvp(u.ctypes.data_as(dp), v.ctypes.data_as(dp), C.c_double(t), p.ctypes.data_as(dp), garr,
                         self.tabw.ctypes.data_as(dp) if self.tabw.size else None,
                         self.tabs.ctypes.data_as(C.POINTER(C.c_int)) if self.tabs.size else None,
                         self.box.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(dp))
        return out

    def unpack(self, u, t, p=None):
        dp, p, garr = self._common(t, p)
        u = np.ascontiguousarray(u, dtype=np.float64)
        nodes = int(np.prod([ax.n for ax in self.prog.axes]))
        out = np.zeros(nodes * len(self.prog.ilo))
        self.lib.emu_unpack(u.ctypes.data_as(dp), C.c_double(t), p.ctypes.data_as(dp), garr,
                            self.tabw.ctypes.data_as(dp) if self.tabw.size else None,
                            self.tabs.ctypes.data_as(C.POINTER(C.c_int)) if self.tabs.size else None, out.ctypes.data_as(dp))
        return out

    def solve(self, u0, alg, t0, t1, dt0=0.0, adaptive=True, abstol=1e-6, reltol=1e-3, saveat=(), maxiters=10 ** 6, p=None):
        """The persistent solver kernel (EmuKernel(..., solve=True)): returns (u(t1), saved states, stats dict)."""
        dp, p, garr = self._common(t0, p)
        u = np.array(u0, dtype=np.float64)
        n = u.size
        work = np.zeros(9 * n)
        sv = np.ascontiguousarray(saveat, dtype=np.float64)
        save = np.zeros(max(1, len(sv)) * n)
        out = np.zeros(7)
        self.lib.emu_solve(u.ctypes.data_as(dp), work.ctypes.data_as(dp), save.ctypes.data_as(dp),
                           sv.ctypes.data_as(dp) if len(sv) else None, len(sv), {"euler": 1, "ssprk33": 2, "rk4": 3, "tsit5": 4}[alg],
                           int(adaptive), C.c_double(t0), C.c_double(t1), C.c_double(dt0), C.c_double(abstol), C.c_double(reltol),
                           C.c_longlong(maxiters), C.c_longlong(n), p.ctypes.data_as(dp), garr,
                           self.tabw.ctypes.data_as(dp) if self.tabw.size else None,
                           self.tabs.ctypes.data_as(C.POINTER(C.c_int)) if self.tabs.size else None,
                           self.box.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(dp))
        stats = dict(t_final=out[0], dt_last=out[1], nf=int(out[2]), naccept=int(out[3]), nreject=int(out[4]), retcode=int(out[5]),
                     nsaved=int(out[6]))
        return u, save.reshape(-1, n)[:len(sv)], stats
