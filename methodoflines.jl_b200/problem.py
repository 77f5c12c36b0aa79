"""discretize / ODEProblem / solve — host layer over libmol_cuda.so.

Mirrors the reference's seam #2: `discretize(pdesys, disc)` returns an ODEProblem whose
in-place RHS `f(du, u, p, t)` is GPU-backed (cf. SciMLBase.discretize override for StaggeredGrid,
src/discretization/staggered_discretize.jl:1-29), and `solve(prob, alg; abstol, reltol, dt,
adaptive, saveat)` runs the explicit RK loop on the device (OrdinaryDiffEq call sites:
test/Diffusion/MOL_1D_Linear_Diffusion.jl:73, benchmark/weno/suite.jl:50-54).

torch is used only as the owner of device memory and streams; everything numeric goes through
the C ABI with raw pointers.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .interface import CudaStencilDiscretization
from .lowering import lower, StencilProgram


class Tsit5:
    name = "tsit5"
    adaptive = True


class SSPRK33:
    name = "ssprk33"
    adaptive = False


class Euler:
    name = "euler"
    adaptive = False


class RK4:
    name = "rk4"
    adaptive = False


def symbolic_discretize(pdesys, disc) -> StencilProgram:
    """The stencil program (text IR + layout) without touching a GPU."""
    return lower(pdesys, disc)


class ODEProblem:
    """GPU-backed ODEProblem: `f(du, u, p, t)` evaluates the semi-discrete RHS on the device."""

    def __init__(self, program: StencilProgram, device: int = 0):
        self.program = program
        self.plan = capi.Plan(program.text, device)
        self.tspan = program.tspan
        self.p = program.pvals.copy()
        self.device = device

    @property
    def u0(self):
        """Initial condition at the unknown nodes (generate_ic_defaults.jl:11-21), evaluated on first use."""
        return self.program.u0

    def f(self, du, u, p, t, stream=None):
        """In-place RHS on torch CUDA tensors (float64, contiguous, length = state_len)."""
        import torch
        assert du.is_cuda and u.is_cuda and du.dtype == torch.float64 and u.dtype == torch.float64
        assert du.numel() == self.plan.state_len and u.numel() == self.plan.state_len
        st = torch.cuda.current_stream(u.device).cuda_stream if stream is None else stream
        self.plan.rhs(du.data_ptr(), u.data_ptr(), t, self.p if p is None else p, st)
        return None

    def jvp(self, jv, u, v, p, t, stream=None):
        """jv = (d f / d u)(u, p, t) v on torch CUDA tensors (mol_jvp)."""
        import torch
        st = torch.cuda.current_stream(u.device).cuda_stream if stream is None else stream
        self.plan.jvp(jv.data_ptr(), u.data_ptr(), v.data_ptr(), t, self.p if p is None else p, st)
        return None

    def rhs_host(self, u_host, t, p=None, nchunks=0):
        """Reference-facing call with HOST buffers (mol_rhs_host): pinned staging, chunked H2D / sweep / D2H
        pipeline inside the library.  Returns a fresh host array."""
        import torch
        n = self.plan.state_len
        if getattr(self, "_pin", None) is None:
            self._pin = (torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory())
        hu, hdu = self._pin
        hu.numpy()[:] = np.asarray(u_host, dtype=np.float64).reshape(-1)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            self.plan.rhs_host(hdu.data_ptr(), hu.data_ptr(), t, self.p if p is None else p, nchunks, st)
            torch.cuda.current_stream().synchronize()
        return hdu.numpy().copy()


class ODESolution:
    """PDETimeSeriesSolution mirror (src/interface/solution/timedep.jl:19-93).

    sol.t / sol[t]      saved times
    sol.u               saved flat unknown vectors (the ODE solution; host copies)
    sol[u(t,x)]         the dependent variable on the WHOLE grid, time axis first: shape (nt, n1, n2, ..) with the
                        boundary nodes rebuilt from the boundary conditions (`observed`) and invalid corner nodes 0 --
                        unpacked on the device by mol_unpack from the saved states (one D2H copy of the result)
    sol[x]              the grid of an independent variable (with domains joined by interfaces: the variable's own grid,
                        and sol[u1(t,x1)] its own node range -- test/Diffusion/MOL_1D_Linear_Diffusion.jl:919-929)
    sol.interior(u)     unknown nodes only, shape (nt, m1, m2, ..)
    """

    def __init__(self, prob, t, u, stats, retcode, save_dev=None):
        self.prob, self.t, self.u, self.stats, self.retcode = prob, np.asarray(t), u, stats, retcode
        self._save_dev = save_dev          # (nt, state_len) torch tensor, still on the device
        self._full = None

    def _varindex(self, key):
        P = self.prob.program
        name = str(getattr(key, "func", key))
        return P.var_names.index(name) if name in P.var_names else None

    def interior(self, key):
        P = self.prob.program
        v = self._varindex(key)
        if v is None:
            raise KeyError(key)
        o, shp = P.offsets[v], P.shapes[v]
        size = int(np.prod(shp))
        return np.stack([np.asarray(uk[o:o + size]).reshape(shp, order="F") for uk in self.u])

    def _unpack(self):
        if self._full is None:
            import torch
            P, plan = self.prob.program, self.prob.plan
            shape = plan.grid_shape(len(P.axes))
            nodes = int(np.prod(shape))
            dev = torch.device("cuda", self.prob.device)
            sv = self._save_dev if self._save_dev is not None else torch.from_numpy(np.stack(self.u)).to(dev)
            full = torch.empty((len(self.t), plan.nvar, nodes), dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                plan.unpack(full.data_ptr(), sv.data_ptr(), self.t, self.prob.p if len(self.prob.p) else None,
                            torch.cuda.current_stream(dev).cuda_stream)
                torch.cuda.current_stream(dev).synchronize()
            self._full = (full.cpu().numpy(), shape)
        return self._full

    def __getitem__(self, key):
        P = self.prob.program
        v = self._varindex(key)
        if v is not None:
            full, shape = self._unpack()
            if P.segments is not None:          # domains joined by interfaces: the variable's own node range of the chart
                seg = P.segments[v]
                out = full[:, v, seg["off"]:seg["off"] + seg["n"]]
                return out[:, 0] if str(seg["sym"]).startswith("__point_") else out      # a variable of t alone: sol[v(t)]
            return full[:, v, :].reshape((len(self.t),) + tuple(reversed(shape))).transpose(
                (0,) + tuple(range(len(shape), 0, -1)))
        for seg in (P.segments or []):
            if key == seg["sym"] or str(key) == str(seg["sym"]):
                return np.array(seg["x"], dtype=float)
        for ax in P.axes:
            if key == ax.sym or str(key) == str(ax.sym):
                return ax.x.copy()
        if str(key) == str(getattr(P, "time", "t")):
            return self.t
        raise KeyError(key)


def discretize(pdesys, disc) -> ODEProblem:
    strat = disc.disc_strategy
    if not isinstance(strat, CudaStencilDiscretization):
        raise TypeError("this backend implements CudaStencilDiscretization only; the reference's Scalarized/Array "
                        "strategies stay in MethodOfLines.jl (interface_errors, MOL_discretization.jl:14-22)")
    return ODEProblem(lower(pdesys, disc), strat.device)


def solve(prob: ODEProblem, alg=None, *, abstol=1e-6, reltol=1e-3, dt=None, adaptive=None, saveat=None,
          maxiters=10 ** 6, save_everystep=False):
    """Explicit RK solve on the device; returns ODESolution with host copies of the saved states."""
    import torch
    alg = alg or Tsit5()
    adaptive = alg.adaptive if adaptive is None else adaptive
    dev = torch.device("cuda", prob.device)
    t0, t1 = prob.tspan
    u = torch.from_numpy(prob.u0).to(dev)
    if saveat is None:
        ts = np.array([t0, t1])
    elif np.isscalar(saveat):
        ts = np.arange(t0, t1 + 0.5 * saveat, saveat)
        ts = ts[ts <= t1 + 1e-12]
    else:
        ts = np.asarray(saveat, dtype=float)
    save = torch.zeros((len(ts), prob.plan.state_len), dtype=torch.float64, device=dev)
    rk = capi.RK(prob.plan, alg.name, abstol, reltol)
    rk.set_params(prob.p) if len(prob.p) else None
    st = rk.solve(u.data_ptr(), t0, t1, 0.0 if dt is None else dt, adaptive, ts, save.data_ptr(), maxiters,
                  torch.cuda.current_stream(dev).cuda_stream)
    rk.close()
    us = save.cpu().numpy()
    stats = dict(nf=st.nf, naccept=st.naccept, nreject=st.nreject)
    return ODESolution(prob, ts, [us[k] for k in range(len(ts))], stats,
                       {0: "Success", 1: "MaxIters", 2: "Unstable"}.get(st.retcode, "Failure"), save_dev=save)
