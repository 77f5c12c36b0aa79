"""Oracle: full-grid NumPy restatement of the reference's scalarized discretization.

Test infrastructure only (see oracle/__init__).  Input: a duck-typed PDESystem /
MOLFiniteDifference pair (attributes eqs, bcs, domains, ivs, dvs, ps / dxs, time,
approx_order, advection_scheme).  Output: `OracleProblem` with `u0`, `rhs(u, t)`.

Architecture (deliberately unlike the GPU path): every discretised variable is
an array over ALL grid nodes (discretize_vars.jl:256-273); each RHS evaluation
  1. scatters the interior unknowns into the full arrays,
  2. fills boundary nodes by solving the BC equations   (generate_bc_eqs.jl:313-328,238-311),
     periodic aliases (:35-58), extrapolation pads (:336-392); corners stay 0 (:396-416),
  3. evaluates each PDE on its interior box with per-node stencil rows chosen exactly
     as the reference does per point:
       centered   centered_difference.jl:5-57
       upwind     upwind_difference.jl:1-28,131-199  (incl. the NU positive-wind index quirk)
       WENO       function_scheme.jl:1-76 + WENO.jl / nonuniform_weno.jl
       nonlinlap  nonlinear_laplacian.jl:28-103 + half_offset_centred_difference.jl:9-69
       spherical  spherical_laplacian.jl:10-42
     with rule precedence of generate_finite_difference_rules.jl:47-116.
Interior boxes follow interior_map.jl:1-10,89-115,117-139.
"""
from fractions import Fraction
import math

import numpy as np
import scipy.sparse as sps
import sympy as sp

from . import operators as ops
from . import weno as wk
from .evalexpr import evaluate, evaluate_abs
from .fornberg import calculate_weights


# ----------------------------------------------------------------------------- grids
def _rat(v):
    f = Fraction(v).limit_denominator(1 << 20)
    return f if float(f) == float(v) else Fraction(v)


def range_values(a, dx, n):
    """Values of the Julia range a:dx:b.  Julia's StepRangeLen uses twice-precision
    arithmetic on rationalised endpoints; restated as exact rational arithmetic rounded once."""
    ra, rdx = _rat(a), _rat(dx)
    return np.array([float(ra + i * rdx) for i in range(n)])


def make_grid(lo, hi, dxspec, edge=False):
    """discretize_space / prepare_dx / generate_grid -- discretize_vars.jl:219-233,283-285,359-390.
    Returns (nodes, dx or None): dx is a float iff the grid is a uniform range.
    edge=True (EdgeAlignedGrid): the nodes are the cell centres of the centre-aligned axis plus one node half a step
    outside each end, so that the domain boundaries sit half-way between the first / last two nodes."""
    if isinstance(dxspec, (int, np.integer)) and not isinstance(dxspec, bool):
        n = int(dxspec)
        dx = (hi - lo) / n if edge else (hi - lo) / (n - 1)          # prepare_dx(::Integer, ...)
        if edge:
            n = n + 1
        g, dxu = range_values(lo, dx, n), float(dx)
    elif np.ndim(dxspec) > 0:
        g = np.asarray(dxspec, dtype=float)
        if g[-1] != hi:
            g = np.append(g, hi)
        dxu = None
    else:
        dx = float(dxspec)
        n = int(math.floor((hi - lo) / dx + 1e-9)) + 1
        g = range_values(lo, dx, n)
        if abs(g[-1] - hi) > 1e-12 * max(1.0, abs(hi)):
            g, dxu = np.append(g, hi), None
        else:
            dxu = dx
    if not edge:
        return g, dxu
    if dxu is not None:                                               # (lo - dx/2):dx:(hi + dx/2)
        return range_values(lo - dxu / 2, dxu, len(g) + 1), dxu
    mid = [(g[i] + g[i + 1]) / 2 for i in range(len(g) - 1)]
    mid = [mid[0] - 2 * (mid[0] - lo)] + mid + [mid[-1] + 2 * (hi - mid[-1])]
    return np.array(mid), None


# ----------------------------------------------------------------------------- parsing
class _Boundary:
    def __init__(self, var, dim, upper, eq):
        self.var, self.dim, self.upper, self.eq = var, dim, upper, eq


def _dv_calls(expr, fn):
    return [a for a in expr.atoms(sp.core.function.AppliedUndef) if a.func == fn]


class OracleProblem:
    def __new__(cls, pdesys=None, disc=None):
        # variables on different domains joined by interfaces have their own restatement (oracle/interface1d.py)
        if cls is OracleProblem and pdesys is not None:
            sa = [[a for a in d.args if a != disc.time] for d in pdesys.dvs]
            if any(q != sa[0] for q in sa):
                from .interface1d import InterfaceOracle1D
                return InterfaceOracle1D(pdesys, disc)
        return super().__new__(cls)

    def __init__(self, pdesys, disc):
        self.sys, self.disc = pdesys, disc
        self.t = disc.time
        self.dvs = list(pdesys.dvs)
        self.funcs = [d.func for d in self.dvs]
        self.nv = len(self.dvs)
        args0 = [a for a in self.dvs[0].args if a != self.t]
        for d in self.dvs:
            assert [a for a in d.args if a != self.t] == args0, \
                "oracle scope: all dependent variables share the same spatial arguments"
        self.xs = args0
        self.nd = len(args0)
        dom = {iv.var: (float(iv.lo), float(iv.hi)) for iv in pdesys.domains}
        self.dom = dom
        self.tspan = dom[self.t]
        self.params = [p for p, _ in pdesys.ps]
        self.pvals = np.array([v for _, v in pdesys.ps], dtype=float)
        self.edge = type(disc.grid_align).__name__ == "EdgeAlignedGrid"
        self.grid, self.dx = [], []
        for x in self.xs:
            g, dx = make_grid(dom[x][0], dom[x][1], disc.dxs[x], self.edge)
            self.grid.append(g)
            self.dx.append(dx)
        self.n = [len(g) for g in self.grid]
        self.weno = type(disc.advection_scheme).__name__ == "WENOScheme"
        self.weno_eps = getattr(disc.advection_scheme, "epsilon", None)
        self.upwind_order = getattr(disc.advection_scheme, "order", 1)
        self._parse_bcs()
        if self.edge:
            assert not self.weno, "oracle scope: edge-aligned grids with centered/upwind schemes"
        self._orders()
        self.dd = [ops.differential_discretizer(self.grid[j], self.dx[j], self.orders[j],
                                                disc.approx_order, self.upwind_order, self.weno)
                   for j in range(self.nd)]
        self._interiors()
        self._initial()

    # -- boundary conditions -------------------------------------------------------------
    def _parse_bcs(self):
        nv, nd = self.nv, self.nd
        self.periodic = [[False] * nd for _ in range(nv)]
        self.bounds = [[[None, None] for _ in range(nd)] for _ in range(nv)]
        self.bounds_more = {}
        self.ics = [None] * nv
        t0 = self.tspan[0]
        for bc in self.sys.bcs:
            resid = bc.lhs - bc.rhs
            found = False
            for v, (dv, fn) in enumerate(zip(self.dvs, self.funcs)):
                for call in _dv_calls(bc.lhs, fn) + _dv_calls(bc.rhs, fn):
                    num = [(k, a) for k, a in enumerate(call.args) if a.is_number]
                    if not num:
                        continue
                    k, val = num[0]
                    canon = dv.args[k]
                    if canon == self.t:
                        assert bc.lhs == call, "initial condition must read u(t0, x...) ~ expr"
                        self.ics[v] = bc.rhs
                        found = True
                        break
                    j = self.xs.index(canon)
                    lo, hi = self.dom[canon]
                    # periodic: u(.., lo, ..) ~ u(.., hi, ..)
                    if (bc.lhs.func == fn and bc.rhs.func == fn and bc.lhs.is_Function
                            and bc.rhs.is_Function):
                        a, b = float(bc.lhs.args[k]), float(bc.rhs.args[k])
                        if {a, b} == {lo, hi}:
                            self.periodic[v][j] = True
                            found = True
                            break
                    upper = abs(float(val) - hi) < 1e-12 * max(1.0, abs(hi))
                    assert upper or abs(float(val) - lo) < 1e-12 * max(1.0, abs(lo)), bc
                    if self.bounds[v][j][int(upper)] is None:
                        self.bounds[v][j][int(upper)] = _Boundary(v, j, upper, bc)
                    else:                                # several conditions at one end: one clipped node each
                        self.bounds_more.setdefault((v, j, int(upper)), []).append(_Boundary(v, j, upper, bc))
                    found = True
                    break
                if found:
                    break
            assert found, f"could not classify boundary condition {bc}"

    def _orders(self):
        """d_orders: derivative orders per spatial variable over PDEs and BCs."""
        self.orders = [set() for _ in range(self.nd)]
        exprs = [e.lhs - e.rhs for e in self.sys.eqs] + [b.lhs - b.rhs for b in self.sys.bcs]
        for e in exprs:
            for D in e.atoms(sp.Derivative):
                for (var, cnt) in D.variable_count:
                    if var in self.xs:
                        self.orders[self.xs.index(var)].add(int(cnt))
        self.orders = [sorted(o) for o in self.orders]

    # -- interior boxes ------------------------------------------------------------------
    def _eqvar(self, eq):
        for D in (eq.lhs - eq.rhs).atoms(sp.Derivative):
            if D.variables == (self.t,) and D.expr in self.dvs:
                return self.dvs.index(D.expr)
        raise ValueError(f"no time derivative in {eq}")

    def _interiors(self):
        self.eq_of_var = {}
        for eq in self.sys.eqs:
            self.eq_of_var[self._eqvar(eq)] = eq
        assert len(self.eq_of_var) == self.nv
        self.lower, self.upper, self.vlower, self.vupper, self.ext = [], [], [], [], []
        for v in range(self.nv):
            eq = self.eq_of_var[v]
            lo = [0] * self.nd
            up = [0] * self.nd
            for j in range(self.nd):
                if self.periodic[v][j]:
                    lo[j] += 1                       # interface lower clips, upper does not
                else:
                    lo[j] += (self.bounds[v][j][0] is not None) + len(self.bounds_more.get((v, j, 0), []))
                    up[j] += (self.bounds[v][j][1] is not None) + len(self.bounds_more.get((v, j, 1), []))
            self.vlower.append(list(lo))
            self.vupper.append(list(up))
            # calculate_stencil_extents (interior_map.jl:117-139)
            le, ue = [0] * self.nd, [0] * self.nd
            resid = eq.lhs - eq.rhs
            for j, x in enumerate(self.xs):
                eqorders = {int(c) for D in resid.atoms(sp.Derivative)
                            for (var, c) in D.variable_count if var == x}
                for d in eqorders:
                    if d % 2 == 1:
                        e = 2 if (d == 1 and self.weno and self.dx[j] is not None) else 0
                        if not self.periodic[v][j]:
                            le[j] = max(le[j], e)
                            ue[j] = max(ue[j], e)
            self.ext.append((le, ue))
            self.lower.append([max(a, b) for a, b in zip(le, lo)])
            self.upper.append([max(a, b) for a, b in zip(ue, up)])
        # interior node ranges, 1-based inclusive
        self.ilo = [[1 + self.lower[v][j] for j in range(self.nd)] for v in range(self.nv)]
        self.ihi = [[self.n[j] - self.upper[v][j] for j in range(self.nd)] for v in range(self.nv)]
        self.ishape = [tuple(self.ihi[v][j] - self.ilo[v][j] + 1 for j in range(self.nd))
                       for v in range(self.nv)]
        self.sizes = [int(np.prod(s)) for s in self.ishape]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(int)
        self.nstate = int(self.offsets[-1])

    def _islice(self, v):
        return tuple(slice(self.ilo[v][j] - 1, self.ihi[v][j]) for j in range(self.nd))

    def _coords(self, v, sl=None):
        sl = sl if sl is not None else self._islice(v)
        out = []
        for j in range(self.nd):
            shape = [1] * self.nd
            g = self.grid[j][sl[j]]
            shape[j] = len(g)
            out.append(g.reshape(shape))
        return out

    def _env(self, coords, t, p, dvvals=None):
        env = {x: c for x, c in zip(self.xs, coords)}
        env[self.t] = t
        env.update({s: float(v) for s, v in zip(self.params, p)})
        if dvvals is not None:
            env.update({dv: val for dv, val in zip(self.dvs, dvvals)})
        return env

    def _initial(self):
        u0 = np.zeros(self.nstate)
        for v in range(self.nv):
            ic = self.ics[v]
            assert ic is not None, f"missing initial condition for {self.dvs[v]}"
            val = evaluate(ic, self._env(self._coords(v), self.tspan[0], self.pvals))
            val = np.broadcast_to(np.asarray(val, dtype=float), self.ishape[v])
            u0[self.offsets[v]:self.offsets[v + 1]] = val.ravel(order="F")
        self.u0 = u0

    # -- state <-> full arrays -----------------------------------------------------------
    def unpack(self, u):
        full = []
        for v in range(self.nv):
            U = np.zeros(self.n)
            U[self._islice(v)] = u[self.offsets[v]:self.offsets[v + 1]].reshape(self.ishape[v], order="F")
            full.append(U)
        return full

    def full_state(self, u, t, p=None):
        """All node values incl. boundary nodes (what sol[u(t,x)] shows per time level)."""
        p = self.pvals if p is None else np.asarray(p, dtype=float)
        full = self.unpack(np.asarray(u, dtype=float))
        self._fill_boundaries(full, t, p)
        return full

    # -- per-node stencil rows (1-based node numbers, as in the reference) -----------------
    def _wrap(self, i, n):
        if i <= 1:
            return i + (n - 1)
        if i > n:
            return i - (n - 1)
        return i

    def centered_row(self, D, i, n, haslower, hasupper):
        """central_difference_weights_and_stencil — centered_difference.jl:5-57."""
        bpc, bsl, L = D.boundary_point_count, D.boundary_stencil_length, D.stencil_length
        if D.uniform:
            if i <= bpc and not haslower:
                return D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)]
            if i > n - bpc and not hasupper:
                return D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)]
            taps = [i + k for k in range(-(L // 2), L // 2 + 1)]
            if haslower or hasupper:
                taps = [self._wrap(tp, n) for tp in taps]
            return D.stencil_coefs, taps
        assert not (haslower or hasupper), "interfaces unsupported on non-uniform centered"
        if i <= bpc:
            return D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)]
        if i > n - bpc:
            return D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)]
        return D.stencil_coefs[i - bpc - 1], [i + k for k in range(-(L // 2), L // 2 + 1)]

    def upwind_row(self, D, i, n, ispositive, haslower, hasupper, grid=None):
        """_upwind_difference — upwind_difference.jl:1-28 (uniform), :131-162 (NU); on a non-uniform grid with a periodic
        wrap, :85-129: where the one-sided stencil crosses the seam, or at the node the shifted table has no row for,
        the weights are computed on the spot from the chart coordinates of the raw taps (period-shifted, :60-72)."""
        L, bsl = D.stencil_length, D.boundary_stencil_length
        wrap = (lambda tp: self._wrap(tp, n)) if (haslower or hasupper) else (lambda tp: tp)
        if D.uniform:
            if not ispositive:
                if i > n - D.boundary_point_count and not hasupper:
                    return D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)]
                return D.stencil_coefs, [wrap(i + k) for k in range(L)]
            if i <= D.offside and not haslower:
                return D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)]
            return D.stencil_coefs, [wrap(i + k) for k in range(-L + 1, 1)]
        if haslower or hasupper:
            assert haslower and hasupper and grid is not None, "oracle scope: periodic wrap (interfaces: oracle/interface1d.py)"
            raw = [i + k for k in (range(L) if not ispositive else range(-L + 1, 1))]
            if any(r < 1 or r > n for r in raw) or (i == n if ispositive else i == 1):
                Lp = grid[-1] - grid[0]
                taps = [self._wrap(r, n) for r in raw]
                xx = [grid[tp - 1] + (Lp if r > tp else -Lp if r < tp else 0.0) for r, tp in zip(raw, taps)]
                return calculate_weights(D.derivative_order, grid[i - 1], xx), taps
        if not ispositive:
            if i > n - D.boundary_point_count:
                return D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)]
            return D.stencil_coefs[i - 1], [i + k for k in range(L)]
        # offside was reset to 0 in the NU constructor: the low branch is unreachable and
        # the row index is NOT shifted (reference quirk)
        return D.stencil_coefs[i - D.offside - 1], [i + k for k in range(-L + 1, 1)]

    def half_row(self, D, i, n, haslower, hasupper, length=0):
        """get_half_offset_weights_and_stencil — half_offset_centred_difference.jl:9-69."""
        ln = n if length == 0 else length
        bpc, bsl, L = D.boundary_point_count, D.boundary_stencil_length, D.stencil_length
        if not D.uniform:
            assert not (haslower or hasupper)
        if i <= bpc and not haslower:
            return D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)]
        if i > ln - bpc and not hasupper:
            return D.high_boundary_coefs[ln - i - 1], [ln - bsl + 1 + k for k in range(bsl)]
        w = D.stencil_coefs if D.uniform else D.stencil_coefs[i - bpc - 1]
        return w, [i + k for k in range(1 - L // 2, L // 2 + 1)]

    # -- sparse application ----------------------------------------------------------------
    def _rows_matrix(self, rows, ncols):
        ri, ci, vv = [], [], []
        for r, (w, taps) in enumerate(rows):
            for wk_, tp in zip(w, taps):
                assert 1 <= tp <= ncols, (tp, ncols)
                ri.append(r)
                ci.append(tp - 1)
                vv.append(wk_)
        return sps.csr_matrix((vv, (ri, ci)), shape=(len(rows), ncols))

    @staticmethod
    def _apply(M, U, axis):
        Um = np.moveaxis(U, axis, 0)
        sh = Um.shape
        R = M @ Um.reshape(sh[0], -1)
        return np.moveaxis(R.reshape((M.shape[0],) + sh[1:]), 0, axis)

    _absmode = False     # rhs_termscale(): every stencil sum becomes sum |w_k| |u_k|

    def _applyd(self, M, U, axis):
        if self._absmode:
            return self._apply(abs(M), np.abs(U), axis)
        return self._apply(M, U, axis)

    def _restrict_other(self, A, v, axis):
        sl = list(self._islice(v))
        sl[axis] = slice(None)
        return A[tuple(sl)]

    # -- derivative evaluations on the interior box of equation-variable `ev` ---------------
    def d_centered(self, full, u, j, d, ev):
        n = self.n[j]
        D = self.dd[j].map[d]
        per = self.periodic[u][j]
        M = self._cached_matrix(("c", u, j, d, ev), lambda: [self.centered_row(D, i, n, per, per)
                                                             for i in range(self.ilo[ev][j], self.ihi[ev][j] + 1)], n)
        return self._restrict_other(self._applyd(M, full[u], j), ev, j)

    def _cached_matrix(self, key, rows_fn, n):
        """The stencil rows do not depend on the state: one sparse matrix per (operator, variable, dimension, equation)."""
        cache = self.__dict__.setdefault("_row_cache", {})
        if key not in cache:
            cache[key] = self._rows_matrix(rows_fn(), n)
        return cache[key]

    def d_mixed(self, full, u, jx, jy, ev):
        """mixed_central_difference — 2nd_order_mixed_deriv.jl:5-22: sum over the taps of the centred first-derivative
        rows of both dimensions, wx wy u[II + xoffset + yoffset].  Taps in the corners of the grid read the corner nodes,
        which are 0 (generate_corner_eqs!, generate_bc_eqs.jl:396-416) -- as they are in the full arrays here."""
        out = full[u]
        for j in (jx, jy):
            n, D, per = self.n[j], self.dd[j].map[1], self.periodic[u][j]
            rows = [self.centered_row(D, i, n, per, per) for i in range(self.ilo[ev][j], self.ihi[ev][j] + 1)]
            out = self._applyd(self._rows_matrix(rows, n), np.abs(out) if (self._absmode and j == jy) else out, j)
        sl = [slice(None)] * self.nd
        for j in range(self.nd):
            if j not in (jx, jy):
                sl[j] = self._islice(ev)[j]
        return out[tuple(sl)]

    def d_upwind(self, full, u, j, d, ev, ispositive):
        n = self.n[j]
        D = (self.dd[j].windneg if ispositive else self.dd[j].windpos)[d]
        per = self.periodic[u][j]
        M = self._cached_matrix(("w", u, j, d, ev, ispositive),
                                lambda: [self.upwind_row(D, i, n, ispositive, per, per, self.grid[j])
                                         for i in range(self.ilo[ev][j], self.ihi[ev][j] + 1)], n)
        return self._restrict_other(self._applyd(M, full[u], j), ev, j)

    def d_weno(self, full, u, j, ev):
        """function_scheme — function_scheme.jl:1-76.  The tap / target / coordinate choice per node does not depend on the
        state: it is made once per (variable, dimension, equation) and the kernel is then evaluated for all nodes that
        share a reconstruction target at once."""
        n = self.n[j]
        per = self.periodic[u][j]
        g = self.grid[j]
        uniform = self.dx[j] is not None
        key = ("weno", u, j, ev)
        cache = self.__dict__.setdefault("_weno_cache", {})
        if key not in cache:
            groups = {}
            for r, i in enumerate(range(self.ilo[ev][j], self.ihi[ev][j] + 1)):
                if i <= 2 and not per:
                    T, raw = i, [1 + k for k in range(5)]
                    taps = raw
                elif i > n - 2 and not per:
                    T, raw = 5 - (n - i), [n - 4 + k for k in range(5)]
                    taps = raw
                else:
                    T, raw = 3, [i + k for k in range(-2, 3)]
                    taps = [self._wrap(tp, n) for tp in raw] if per else raw
                if uniform:
                    assert T == 3, "uniform WENO is only defined on the interior (extent 2)"
                    xx = None
                elif per:
                    # bcoord: exact chart coordinates across the periodic seam (interface_boundary.jl:120-153)
                    Lp = g[-1] - g[0]
                    xx = [g[tp - 1] - Lp if rw <= 1 and rw != tp else
                          (g[tp - 1] + Lp if rw > n else g[tp - 1]) for rw, tp in zip(raw, taps)]
                else:
                    xx = [g[tp - 1] for tp in taps]
                grp = groups.setdefault(T, ([], [], []))
                grp[0].append(r); grp[1].append([tp - 1 for tp in taps]); grp[2].append(xx)
            cache[key] = {T: (np.array(a), np.array(b), None if c[0] is None else np.array(c, dtype=float))
                          for T, (a, b, c) in groups.items()}
        Um = np.moveaxis(self._restrict_other(full[u], ev, j), j, 0)
        out = np.zeros((self.ishape[ev][j],) + Um.shape[1:])
        bshape = (-1,) + (1,) * (Um.ndim - 1)
        for T, (rows, taps, xx) in cache[key].items():
            uu = [Um[taps[:, k]] for k in range(5)]
            if xx is None:
                out[rows] = wk.weno_f_uniform(uu, self.weno_eps, self.dx[j])
            else:
                out[rows] = wk.weno_f_nonuniform_core(uu, self.weno_eps, [xx[:, k].reshape(bshape) for k in range(5)], T)
        if self._absmode:
            out = np.abs(out)
        return np.moveaxis(out, 0, j)

    def nonlinlap(self, full, t, p, inner, u, j, ev):
        """cartesian_nonlinear_laplacian — nonlinear_laplacian.jl:28-103.
        `inner` = expression multiplying Dx(u) inside the outer derivative."""
        n = self.n[j]
        dd = self.dd[j]
        per = self.periodic[u][j]
        x = self.xs[j]
        g = self.grid[j]
        # half points m = 1..n-1 (m sits between nodes m and m+1); with periodic wrap also m = n -> 1
        hp = list(range(1, n)) if not per else list(range(0, n + 1))
        interp_rows, deriv_rows = [], []
        wrapf = (lambda tp: self._wrap(tp, n)) if per else (lambda tp: tp)
        for m in hp:
            w, taps = self.half_row(dd.interp, m, n, per, per)
            interp_rows.append((w, [wrapf(tp) for tp in taps]))
            w, taps = self.half_row(dd.half_inner[1], m, n, per, per)
            deriv_rows.append((w, [wrapf(tp) for tp in taps]))
        Mi = self._rows_matrix(interp_rows, n)
        Md = self._rows_matrix(deriv_rows, n)
        ui = [self._restrict_other(self._apply(Mi, full[v], j), ev, j) for v in range(self.nv)]
        du = self._restrict_other(self._applyd(Md, full[u], j), ev, j)
        xh = Mi @ g
        coords = self._coords(ev)
        shape = [1] * self.nd
        shape[j] = len(hp)
        coords[j] = xh.reshape(shape)
        a = evaluate(inner, self._env(coords, t, p, ui))
        if self._absmode:
            a = np.abs(a)
        flux = np.broadcast_to(a, du.shape) * du
        # outer half-offset difference at II - 1 on the clipped grid (length n-1)
        outer_rows = []
        hp_index = {m: k for k, m in enumerate(hp)}
        for i in range(self.ilo[ev][j], self.ihi[ev][j] + 1):
            w, taps = self.half_row(dd.half_outer, i - 1, n, per, per, length=n - 1)
            outer_rows.append((w, [hp_index[tp] + 1 for tp in taps]))
        Mo = self._rows_matrix(outer_rows, len(hp))
        return self._applyd(Mo, flux, j)

    # -- boundary fill ---------------------------------------------------------------------
    def _edge_slices(self, v, j, upper, node=None):
        """Index tuple of the boundary-face nodes owned by `edge()` (generate_bc_eqs.jl:5-22):
        the interior's extent in the other dims, cast onto node 1 / n (or `node`) in dim j."""
        sl = list(self._islice(v))
        idx = (self.n[j] if upper else 1) if node is None else node
        sl[j] = slice(idx - 1, idx)
        return tuple(sl)

    def _fill_boundaries(self, full, t, p):
        for v in range(self.nv):
            U = full[v]
            # 1. periodic alias u[1] ~ u[n]
            for j in range(self.nd):
                if self.periodic[v][j]:
                    sl1 = list(self._islice(v)); sl1[j] = slice(0, 1)
                    sln = list(self._islice(v)); sln[j] = slice(self.n[j] - 1, self.n[j])
                    U[tuple(sl1)] = U[tuple(sln)]
        # 2. truncating boundaries: affine solve for the edge node
        solved_pads = set()
        for v in range(self.nv):
            for j in range(self.nd):
                for side in (0, 1):
                    b = self.bounds[v][j][side]
                    if b is None:
                        continue
                    bs = [b] + self.bounds_more.get((v, j, side), [])
                    pads = self._pads(v, j, bool(side))
                    hd = self.__dict__.setdefault("_has_deriv", {})
                    if (v, j, side) not in hd:
                        hd[(v, j, side)] = any((q.eq.lhs - q.eq.rhs).atoms(sp.Derivative) for q in bs)
                    has_deriv = hd[(v, j, side)]
                    if pads and has_deriv:
                        # a derivative condition next to extrapolation pads (uniform WENO + Neumann / Robin): the one-sided
                        # row reads the pad node and the pad's extrapolation row reads the edge node; ModelingToolkit
                        # solves the two algebraic equations together, and so does this
                        self._solve_bc_set(full, bs, t, p, pads)
                        solved_pads.update((v, j, nd_) for nd_ in pads)
                    elif len(bs) > 1:
                        self._solve_bc_set(full, bs, t, p)
                    else:
                        self._solve_bc(full, b, t, p)
        # 3. extrapolation pads (generate_extrap_eqs! — generate_bc_eqs.jl:336-392)
        # An edge node belongs to exactly one dimension's edge set (its other indices lie in the
        # interior), so the reference's overlap averaging (:386-389) never triggers here.
        for v in range(self.nv):
            le, ue = self.ext[v]
            for j in range(self.nd):
                if self.periodic[v][j]:
                    continue
                n = self.n[j]
                B = self.dd[j].boundary
                for upper, e, vl in ((False, le[j], self.vlower[v][j]), (True, ue[j], self.vupper[v][j])):
                    ninterp = e - vl
                    while ninterp >= vl:
                        node = (n - ninterp) if upper else (1 + ninterp)
                        ninterp -= 1
                        if self.ilo[v][j] <= node <= self.ihi[v][j] or (v, j, node) in solved_pads:
                            continue                      # interior nodes get no pad equation
                        if vl == 0:
                            raise NotImplementedError(
                                "oracle scope: extrapolation pads coupled to an unconstrained boundary node")
                        w, taps = self.centered_row(B, node, n, False, False)
                        sl = self._edge_slices(v, j, upper, node)
                        val = np.zeros(full[v][sl].shape)
                        for wk_, tp in zip(w, taps):
                            s2 = list(sl); s2[j] = slice(tp - 1, tp)
                            val = val + wk_ * full[v][tuple(s2)]
                        full[v][sl] = val

    def _solve_bc(self, full, b, t, p):
        """Boundary node from the boundary condition (generate_bc_eqs.jl:313-328).  Centre-aligned grid
        (boundary_value_maps :238-311): u(t, x_b) -> the edge node, Dx^d u(t, x_b) -> the one-sided row of the centred
        operator at the edge node.  Edge-aligned grid (:79-161): the boundary lies half-way between the first / last
        two nodes; u(t, x_b) -> the interpolation row (CompleteHalfCenteredDifference(0, max(4, p))) and Dx^d u(t, x_b)
        -> the half-offset derivative row at that half point, II = 1 or len - 1 (newindex(...; shift = true))."""
        cache = self.__dict__.setdefault("_bc_cache", {})
        if id(b) not in cache:
            cache[id(b)] = self._prepare_bc(b)
        resid, recipes, sl, xb, ub = cache[id(b)]
        v, j = b.var, b.dim
        coords = self._coords(v, sl)
        coords[j] = np.full([1] * self.nd, xb)
        env = self._env(coords, t, p)
        env.update({s_: full[w_][sl2] for s_, (w_, sl2) in recipes.items()})
        shape = full[v][sl].shape
        F0 = np.broadcast_to(evaluate(resid, {**env, ub: 0.0}), shape)
        F1 = np.broadcast_to(evaluate(resid, {**env, ub: 1.0}), shape)
        full[v][sl] = -F0 / (F1 - F0)

    def _prepare_bc(self, b):
        """The symbolic part of _solve_bc (independent of the state and of t): the boundary residual with the edge node
        as the symbol `ub` and every other tap as a placeholder symbol bound to (variable, index tuple)."""
        v, j, upper = b.var, b.dim, b.upper
        n = self.n[j]
        x = self.xs[j]
        xb = self.dom[x][1 if upper else 0] if self.edge else (self.grid[j][-1] if upper else self.grid[j][0])
        sl = self._edge_slices(v, j, upper)
        resid = b.eq.lhs - b.eq.rhs
        ub = sp.Symbol("__ub")
        subs = {}
        placeholders = {}
        node = n if upper else 1
        half = n - 1 if upper else 1

        def row_expr(w, taps, w_, tag):
            expr = 0
            for k, (wk_, tp) in enumerate(zip(w, taps)):
                if tp == node and w_ == v:
                    expr = expr + float(wk_) * ub
                else:
                    s_ = sp.Symbol(f"__tap_{w_}_{tag}_{k}")
                    s2 = list(sl); s2[j] = slice(tp - 1, tp)
                    placeholders[s_] = (w_, tuple(s2))
                    expr = expr + float(wk_) * s_
            return expr
        # derivative atoms at the boundary
        for D in resid.atoms(sp.Derivative):
            call = D.expr
            assert call.func in self.funcs, f"unsupported BC derivative {D}"
            w_ = self.funcs.index(call.func)
            (var, cnt), = D.variable_count
            assert var == x, f"BC derivative must be normal to the boundary: {D}"
            if self.edge:
                w, taps = self.half_row(self.dd[j].half_inner[int(cnt)], half, n, False, False)
            else:
                w, taps = self.centered_row(self.dd[j].map[int(cnt)], node, n, False, False)
            subs[D] = row_expr(w, taps, w_, f"d{int(cnt)}")
        resid = resid.xreplace(subs)
        # dependent variables evaluated at the boundary
        for w_, fn in enumerate(self.funcs):
            for call in _dv_calls(resid, fn):
                if self.edge:
                    w, taps = self.half_row(self.dd[j].interp, half, n, False, False)
                    resid = resid.xreplace({call: row_expr(w, taps, w_, "i")})
                elif w_ == v:
                    resid = resid.xreplace({call: ub})
                else:
                    s = sp.Symbol(f"__bv_{w_}")
                    placeholders[s] = (w_, sl)
                    resid = resid.xreplace({call: s})
        return resid, placeholders, sl, xb, ub

    def _pads(self, v, j, upper):
        """Nodes that get an extrapolation equation at one end (generate_extrap_eqs!, generate_bc_eqs.jl:336-392)."""
        le, ue = self.ext[v]
        n = self.n[j]
        e, vl = (ue[j], self.vupper[v][j]) if upper else (le[j], self.vlower[v][j])
        out, ninterp = [], e - vl
        while ninterp >= vl and vl > 0:
            node = (n - ninterp) if upper else (1 + ninterp)
            ninterp -= 1
            if not (self.ilo[v][j] <= node <= self.ihi[v][j]):
                out.append(node)
        return out

    def _solve_bc_set(self, full, bs, t, p, pads=()):
        """m boundary conditions at one end: the m clipped nodes next to that end are the unknowns of the m boundary
        equations, every one of them written at the EDGE node (u(t, x_b) -> u[edge], Dx^d u(t, x_b) -> the one-sided
        row of the centred operator at the edge node: boundary_value_maps, generate_bc_eqs.jl:238-311; interior clipped
        by one node per condition, interior_map.jl:1-10).  Affine system solved per boundary-face point."""
        assert not self.edge, "oracle scope: several conditions per end on centre-aligned grids"
        v, j, upper = bs[0].var, bs[0].dim, bs[0].upper
        n, m = self.n[j], len(bs)
        x = self.xs[j]
        node = n if upper else 1
        nodes = [node - k if upper else node + k for k in range(m)] + list(pads)
        m = len(nodes)
        xb = self.grid[j][node - 1]
        sl = self._edge_slices(v, j, upper)
        ubs = [sp.Symbol(f"__ub{k}") for k in range(m)]
        placeholders, resids = {}, []
        for q, b in enumerate(bs):
            resid = b.eq.lhs - b.eq.rhs
            subs = {}
            for D in resid.atoms(sp.Derivative):
                assert D.expr.func in self.funcs
                w_ = self.funcs.index(D.expr.func)
                (var, cnt), = D.variable_count
                assert var == x
                w, taps = self.centered_row(self.dd[j].map[int(cnt)], node, n, False, False)
                expr = 0
                for k, (wk_, tp) in enumerate(zip(w, taps)):
                    if w_ == v and tp in nodes:
                        expr = expr + float(wk_) * ubs[nodes.index(tp)]
                    else:
                        s_ = sp.Symbol(f"__tap_{q}_{w_}_{int(cnt)}_{k}")
                        s2 = list(sl); s2[j] = slice(tp - 1, tp)
                        placeholders[s_] = full[w_][tuple(s2)]
                        expr = expr + float(wk_) * s_
                subs[D] = expr
            resid = resid.xreplace(subs)
            for w_, fn in enumerate(self.funcs):
                for call in _dv_calls(resid, fn):
                    if w_ == v:
                        resid = resid.xreplace({call: ubs[0]})
                    else:
                        s_ = sp.Symbol(f"__bv_{q}_{w_}")
                        placeholders[s_] = full[w_][sl]
                        resid = resid.xreplace({call: s_})
            resids.append(resid)
        for pad in pads:                      # u[pad] - sum_k w_k u[tap_k] = 0 with the boundary extrapolation row of the pad
            w, taps = self.centered_row(self.dd[j].boundary, pad, n, False, False)
            expr = ubs[nodes.index(pad)]
            for k, (wk_, tp) in enumerate(zip(w, taps)):
                if wk_ == 0.0:
                    continue
                if tp in nodes:
                    expr = expr - float(wk_) * ubs[nodes.index(tp)]
                else:
                    s_ = sp.Symbol(f"__pad_{pad}_{k}")
                    s2 = list(sl); s2[j] = slice(tp - 1, tp)
                    placeholders[s_] = full[v][tuple(s2)]
                    expr = expr - float(wk_) * s_
            resids.append(expr)
        coords = self._coords(v, sl)
        coords[j] = np.full([1] * self.nd, xb)
        env = self._env(coords, t, p)
        env.update(placeholders)
        shape = full[v][sl].shape

        def R(vals):
            e2 = {**env, **{ub: val for ub, val in zip(ubs, vals)}}
            return np.stack([np.broadcast_to(np.asarray(evaluate(r, e2), dtype=float), shape) for r in resids], axis=-1)
        F0 = R([0.0] * m)
        A = np.stack([R([1.0 if i == k else 0.0 for i in range(m)]) - F0 for k in range(m)], axis=-1)   # [..., eq, unknown]
        sol = np.linalg.solve(A, -F0[..., None])[..., 0]
        for k, nd_ in enumerate(nodes):
            s2 = list(sl); s2[j] = slice(nd_ - 1, nd_)
            full[v][tuple(s2)] = sol[..., k]

    # -- term lowering -----------------------------------------------------------------------
    def _lower_term(self, term, ev, ph):
        """Replace derivative structure in one additive term by placeholder symbols bound to thunks
        (full, t, p) -> interior-box array, with the rule precedence of generate_finite_difference_rules.jl.  The
        structure does not depend on the state, so rhs() builds it once per equation."""
        def new(thunk):
            s = sp.Symbol(f"__d{len(ph)}")
            ph[s] = thunk
            return s

        factors = list(sp.Mul.make_args(term))
        # spherical: r^-2 * Dr(r^2 * a * Dr(u))
        for k, fct in enumerate(factors):
            if isinstance(fct, sp.Derivative) and len(fct.variable_count) == 1 and fct.variable_count[0][1] == 1:
                r = fct.variable_count[0][0]
                inner = list(sp.Mul.make_args(fct.expr))
                dus = [q for q in inner if isinstance(q, sp.Derivative) and q.expr in self.dvs
                       and q.variable_count == ((r, 1),)]
                if len(dus) == 1 and r in self.xs:
                    j = self.xs.index(r)
                    u = self.dvs.index(dus[0].expr)
                    rest_in = [q for q in inner if q is not dus[0]]
                    others = factors[:k] + factors[k + 1:]
                    if sp.Pow(r, -2) in others and sp.Pow(r, 2) in rest_in:
                        others.remove(sp.Pow(r, -2))
                        rest_in.remove(sp.Pow(r, 2))
                        a = sp.Mul(*rest_in)
                        val = lambda full, t, p, a=a, u=u, j=j: self._spherical(full, t, p, a, u, j, ev)
                        return sp.Mul(*others) * new(val) if others else new(val)
                    a = sp.Mul(*rest_in)
                    val = lambda full, t, p, a=a, u=u, j=j: self.nonlinlap(full, t, p, a, u, j, ev)
                    return self._lower_generic(sp.Mul(*others), ev, ph) * new(val)
        # upwind: coef * Dx^d(u), d odd, as a direct factor
        for k, fct in enumerate(factors):
            if isinstance(fct, sp.Derivative) and fct.expr in self.dvs and len(fct.variable_count) == 1:
                x, d = fct.variable_count[0]
                d = int(d)
                if x in self.xs and d % 2 == 1 and not (self.weno and d == 1) and len(factors) > 1:
                    j = self.xs.index(x)
                    u = self.dvs.index(fct.expr)
                    coef = sp.Mul(*(factors[:k] + factors[k + 1:]))
                    assert not coef.atoms(sp.Derivative), "derivatives inside an upwind coefficient"
                    bwd = new(lambda full, t, p, u=u, j=j, d=d: self.d_upwind(full, u, j, d, ev, True))
                    fwd = new(lambda full, t, p, u=u, j=j, d=d: self.d_upwind(full, u, j, d, ev, False))
                    return sp.Piecewise((coef * bwd, coef > 0), (coef * fwd, True))
        return self._lower_generic(term, ev, ph)

    def _lower_generic(self, expr, ev, ph):
        subs = {}
        for D in expr.atoms(sp.Derivative):
            if D.expr in self.dvs and len(D.variable_count) == 2 and all(int(c) == 1 for _, c in D.variable_count):
                (xa, _), (xb, _) = D.variable_count
                s = sp.Symbol(f"__d{len(ph)}")
                ph[s] = lambda full, t, p, u=self.dvs.index(D.expr), ja=self.xs.index(xa), jb=self.xs.index(xb): \
                    self.d_mixed(full, u, ja, jb, ev)
                subs[D] = s
                continue
            assert D.expr in self.dvs and len(D.variable_count) == 1, f"unsupported derivative {D}"
            x, d = D.variable_count[0]
            d = int(d)
            j = self.xs.index(x)
            u = self.dvs.index(D.expr)
            if d % 2 == 0:
                arr = lambda full, t, p, u=u, j=j, d=d: self.d_centered(full, u, j, d, ev)
            elif self.weno and d == 1:
                arr = lambda full, t, p, u=u, j=j: self.d_weno(full, u, j, ev)
            else:
                arr = lambda full, t, p, u=u, j=j, d=d: self.d_upwind(full, u, j, d, ev, True)
            s = sp.Symbol(f"__d{len(ph)}")
            ph[s] = arr
            subs[D] = s
        return expr.xreplace(subs)

    def _spherical(self, full, t, p, a, u, j, ev):
        """spherical_diffusion — spherical_laplacian.jl:10-42."""
        r = self.grid[j][self.ilo[ev][j] - 1:self.ihi[ev][j]]
        shape = [1] * self.nd
        shape[j] = len(r)
        rr = r.reshape(shape)
        here = evaluate(a, self._env(self._coords(ev), t, p,
                                     [full[v][self._islice(ev)] for v in range(self.nv)]))
        d2 = self.d_centered(full, u, j, 2, ev)
        d1 = self.d_centered(full, u, j, 1, ev)
        nl = self.nonlinlap(full, t, p, a, u, j, ev)
        near0 = np.abs(rr) <= 1e-6
        with np.errstate(divide="ignore", invalid="ignore"):
            if self._absmode:
                here = np.abs(here)
                reg = here * (d1 / np.abs(rr) + nl)
            else:
                reg = here * (d1 / rr + nl)
        return np.where(near0, 6 * here * d2, reg)

    @staticmethod
    def split_additive(expr):
        """Flatten sums and push numeric coefficients into nested sums (what SymbolicUtils' Add
        canonical form does); products of symbolic factors are NOT distributed."""
        out = []
        for term in sp.Add.make_args(expr):
            c, rest = term.as_coeff_Mul()
            if isinstance(rest, sp.Add):
                out += [c * q for q in OracleProblem.split_additive(rest)]
            else:
                out.append(term)
        return out

    # -- the RHS -------------------------------------------------------------------------
    def rhs_termscale(self, u, t, p=None):
        """Per-unknown sum of the MAGNITUDES of everything that is added up to form du
        (sum_k |w_k||u_k| for stencil rows, |a||b| for products, sum |term| for sums).  This is the
        scale floating-point rounding of one RHS evaluation is proportional to; parity tests
        normalise by it where du itself is small through cancellation (SURVEY §7 hard part 1)."""
        self._absmode = True
        try:
            return self.rhs(u, t, p)
        finally:
            self._absmode = False

    def rhs(self, u, t, p=None):
        p = self.pvals if p is None else np.asarray(p, dtype=float)
        full = self.unpack(np.asarray(u, dtype=float))
        self._fill_boundaries(full, t, p)
        du = np.zeros(self.nstate)
        for ev in range(self.nv):
            cache = self.__dict__.setdefault("_eq_cache", {})
            if ev not in cache:
                eq = self.eq_of_var[ev]
                resid = eq.lhs - eq.rhs                          # cardinalised: lhs - rhs ~ 0
                dt_term = sp.Derivative(self.dvs[ev], self.t)
                cdt = sp.expand(resid).coeff(dt_term)            # c Dt(u) + rest ~ 0: du/dt = -rest / c, terms discretised as written
                rest = sp.expand(resid) - cdt * dt_term if cdt != 1 else resid - dt_term
                assert cdt.is_number and cdt != 0 and not rest.has(dt_term)
                ph = {}
                lowered = sum(self._lower_term(term, ev, ph) for term in self.split_additive(rest))
                cache[ev] = (sp.sympify(lowered), ph, float(cdt))
            lowered, thunks, cdt = cache[ev]
            here = [full[v][self._islice(ev)] for v in range(self.nv)]
            env = self._env(self._coords(ev), t, p, here)
            env.update({s_: th(full, t, p) for s_, th in thunks.items()})
            val = (evaluate_abs if self._absmode else evaluate)(lowered, env)
            scale = (1.0 / abs(float(cdt))) if self._absmode else (-1.0 / float(cdt))
            val = scale * np.broadcast_to(np.asarray(val, dtype=float), self.ishape[ev])
            du[self.offsets[ev]:self.offsets[ev + 1]] = val.ravel(order="F")
        return du
