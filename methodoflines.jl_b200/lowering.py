"""Lowering pass: PDESystem + MOLFiniteDifference  ->  stencil program (text IR for libmol_cuda).

This is the "new lowering pass beside src/array_discretization.jl" of the north star.  Where the
reference's ScalarizedDiscretization emits one symbolic equation per grid point
(src/scalar_discretization.jl:1-64) and ArrayDiscretization one slice equation per
translation-invariant box (src/array_discretization.jl:153-235), this pass emits per equation
ONE pointwise RPN expression over stencil-row *tables*:

  tab   rows indexed by node (or half point): first tap + weights.  A contiguous "core" range
        shares one literal row (the reference's core box, array_discretization.jl:368-420); the
        rest ("frame", non-uniform grids) are explicit rows.  Row selection per node restates
          centered   centered_difference.jl:5-57        upwind  upwind_difference.jl:1-28,131-162
          half-point half_offset_centred_difference.jl:9-69
        with the weight tables of centered_diff_weights.jl / upwind_diff_weights.jl /
        half_offset_weights.jl / extrapolation_weights.jl built from mol_fd_weights (C ABI).
  wtab  WENO5 rows: first tap + reconstruction target (function_scheme.jl:1-34).
  ghost value of a node outside the interior box as  g(t, x) + sum_k a_k * u[interior tap k]:
        Dirichlet data, the solved affine Neumann/Robin condition (generate_bc_eqs.jl:238-328)
        and the order-6 extrapolation pad (:336-392) all reduce to this form.
  eq    du_v/dt as RPN over constants, parameters, t, node coordinates, field values and
        L (linear row), W (WENO), N (nonlinear Laplacian, nonlinear_laplacian.jl:28-103) ops.
        Upwinding is the reference's own ifelse(coef > 0, coef*backward, coef*forward)
        (upwind_difference.jl:192-198) on the cardinalised residual lhs - rhs ~ 0.

Rule precedence follows generate_finite_difference_rules.jl:47-116 (spherical, nonlinear
Laplacian, centered, advection).  Unsupported patterns raise StencilLoweringError (the analogue of
ArrayDiscretizationError under StrictArrayDiscretization, array_discretization.jl:42-59).
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np
import sympy as sp

from . import capi


class StencilLoweringError(NotImplementedError):
    pass


def _hex(v):
    return float(v).hex()


# ---------------------------------------------------------------------------------------- grids
def _rationalize(v):
    f = Fraction(v).limit_denominator(1 << 20)
    return f if float(f) == float(v) else Fraction(v)


def uniform_nodes(a, dx, n):
    """Node values of the range a:dx:b (exact rational arithmetic, rounded once per node)."""
    ra, rd = _rationalize(a), _rationalize(dx)
    den = ra.denominator * rd.denominator // math.gcd(ra.denominator, rd.denominator)
    na, nd = ra.numerator * (den // ra.denominator), rd.numerator * (den // rd.denominator)
    return np.array([(na + k * nd) / den for k in range(n)])       # int / int is correctly rounded, like float(Fraction)


class Axis:
    """One spatial dimension of the DiscreteSpace (discretize_vars.jl:219-251,283-285; generate_grid :359-390).
    edge=True (EdgeAlignedGrid): the nodes are the cell centres of the centre-aligned axis plus one node half a step
    outside each end; `lo`/`hi` stay the domain boundaries (half-way between the first / last two nodes)."""

    def __init__(self, sym, lo, hi, spec, edge=False):
        self.sym, self.lo, self.hi, self.edge = sym, float(lo), float(hi), bool(edge)
        if isinstance(spec, (int, np.integer)) and not isinstance(spec, bool):
            self.n = int(spec) + 1 if edge else int(spec)                 # prepare_dx(::Integer, ...)
            self.dx = (self.hi - self.lo) / (self.n - 1)
            self.x = uniform_nodes(self.lo, self.dx, self.n)
        elif np.ndim(spec) > 0:
            x = np.asarray(spec, dtype=float)
            if x[-1] != self.hi:
                x = np.append(x, self.hi)
            self.x, self.n, self.dx = x, len(x), None
        else:
            dx = float(spec)
            n = int(math.floor((self.hi - self.lo) / dx + 1e-9)) + 1
            x = uniform_nodes(self.lo, dx, n)
            if abs(x[-1] - self.hi) > 1e-12 * max(1.0, abs(self.hi)):
                self.x, self.n, self.dx = np.append(x, self.hi), n + 1, None
            else:
                self.x, self.n, self.dx = x, n, dx
        if edge:
            if self.dx is not None:                                       # (lo - dx/2):dx:(hi + dx/2)
                self.n += 1
                self.x = uniform_nodes(self.lo - self.dx / 2, self.dx, self.n)
            else:
                g = self.x
                mid = [(g[i] + g[i + 1]) / 2 for i in range(len(g) - 1)]
                mid = [mid[0] - 2 * (mid[0] - self.lo)] + mid + [mid[-1] + 2 * (self.hi - mid[-1])]
                self.x, self.n = np.array(mid), len(mid)

    @property
    def uniform(self):
        return self.dx is not None


# ---------------------------------------------------------------------------------------- weights
def _ends(periodic):
    """`periodic` is a bool (both ends wrap) or a pair (lower end open, upper end open): an end is open when a
    periodic / interface boundary sits there, i.e. the scheme never selects a one-sided boundary row at that end
    (haslowerupper, centered_difference.jl:14, upwind_difference.jl:7, function_scheme.jl:4)."""
    if isinstance(periodic, tuple):
        return bool(periodic[0]), bool(periodic[1])
    return bool(periodic), bool(periodic)


class AxisStencils:
    """Stencil rows of one axis.  All `*_row(i)` take and return 1-based node numbers and give
    (first_tap, weights) with RAW (unwrapped) taps; periodic wrap is applied by the kernels."""

    def __init__(self, axis: Axis, approx_order: int, upwind_order: int):
        self.ax, self.p, self.pu = axis, approx_order, upwind_order
        self._c = {}

    def _memo(self, key, fn):
        if key not in self._c:
            self._c[key] = fn()
        return self._c[key]

    # -- centered (CompleteCenteredDifference + central_difference_weights_and_stencil) ----------
    def centered_row(self, d, i, periodic, p=None):
        p = self.p if p is None else p
        n, x = self.ax.n, self.ax.x
        L = d + p - 1 + (d + p) % 2
        bsl, bpc = d + p, L // 2
        lo_open, up_open = _ends(periodic)
        if self.ax.uniform:
            s = 1.0 / self.ax.dx ** d
            if i <= bpc and not lo_open:
                w = self._memo(("cl", d, p, i), lambda: s * capi.fd_weights(d, float(i - 1), np.arange(bsl, dtype=float)))
                return 1, w
            if i > n - bpc and not up_open:
                k = n - i            # mirrored low row k (0-based), reversed, sign (-1)^d
                w = self._memo(("ch", d, p, k), lambda: (s * capi.fd_weights(d, float(k), np.arange(bsl, dtype=float)))[::-1]
                               * (-1.0) ** d)
                return n - bsl + 1, w
            w = self._memo(("ci", d, p), lambda: s * capi.fd_weights(d, 0.0, np.arange(-(L // 2), L // 2 + 1, dtype=float)))
            return i - L // 2, w
        if lo_open or up_open:
            raise StencilLoweringError("periodic/interface boundaries are not supported on non-uniform grids for centered "
                                       "differences (centered_difference.jl:37)")
        dxs = self._memo("dxs", lambda: np.diff(self.ax.x))       # (per axis, not per row)
        if i <= bpc:
            lx = np.concatenate([[0.0], np.cumsum(dxs[:bsl - 1])])
            return 1, capi.fd_weights(d, lx[i - 1], lx)
        if i > n - bpc:
            hx = np.cumsum(dxs[n - 1 - bsl:])
            return n - bsl + 1, capi.fd_weights(d, hx[len(hx) - 1 - (n - i)], hx)
        # interior rows of every node in one library call (centered_diff_weights.jl:94-103)
        W = self._memo(("cnu", d, p), lambda: capi.fd_weights_rows(
            d, x[bpc:n - bpc], np.lib.stride_tricks.sliding_window_view(x, 2 * bpc + 1)))
        return i - bpc, W[i - 1 - bpc]

    # -- upwind (CompleteUpwindDifference + _upwind_difference) ------------------------------------
    def upwind_row(self, d, i, positive, periodic, coord=None):
        """coord(r): chart coordinate of the raw (unwrapped) node r, which may lie past an open end (bcoord,
        interface_boundary.jl:109-153); only needed on non-uniform grids with an open end."""
        n, x = self.ax.n, self.ax.x
        L = d + self.pu
        lo_open, up_open = _ends(periodic)
        if self.ax.uniform:
            s = 1.0 / self.ax.dx ** d
            if not positive:       # forward operator, offside 0, high boundary rows L-1
                if i > n - (L - 1) and not up_open:
                    # REFERENCE QUIRK: the high rows are computed for spots 0,-1,..,-(L-2) on mirrored nodes
                    # 0,-1,..,-(L-1) and the LIST is reversed (upwind_diff_weights.jl:49-77), so node i gets the
                    # row of spot -(L-2-(n-i)).  Exact for first-order upwind (L = 2); mirrored as is otherwise.
                    k = (L - 2) - (n - i)
                    w = self._memo(("uh", d, k), lambda: ((-1.0 / self.ax.dx) ** d)
                                   * capi.fd_weights(d, -float(k), -np.arange(L, dtype=float)))
                    return n - L + 1, w
                w = self._memo(("uf", d), lambda: s * capi.fd_weights(d, 0.0, np.arange(L, dtype=float)))
                return i, w
            off = L - 1            # backward operator, offside d+p-1
            if i <= off and not lo_open:
                w = self._memo(("ul", d, i), lambda: s * capi.fd_weights(d, float(i - 1), np.arange(L, dtype=float)))
                return 1, w
            w = self._memo(("ub", d), lambda: s * capi.fd_weights(d, 0.0, np.arange(L, dtype=float) - off))
            return i - L + 1, w
        if lo_open or up_open:
            # upwind_difference.jl:85-129: where the one-sided stencil leaves the grid through an open end, or at the
            # "blind spot" node the shifted table has no row for, the weights are computed on the spot from the chart
            # coordinates of the raw taps (no index quirk here); elsewhere the standard non-uniform rows below apply.
            raw = list(range(i - L + 1, i + 1)) if positive else list(range(i, i + L))
            crossing = any(r < 1 or r > n for r in raw)
            if crossing or (i == n if positive else i == 1):
                for r in raw:
                    if (r < 1 and not lo_open) or (r > n and not up_open):
                        raise StencilLoweringError(
                            "the upwind stencil extends past a non-interface boundary on a nonuniform grid "
                            "(upwind_difference.jl:111-115)")
                if coord is None:
                    raise StencilLoweringError("chart coordinates are needed for a non-uniform upwind row at an open end")
                return raw[0], capi.fd_weights(d, coord(i), np.array([coord(r) for r in raw], dtype=float))
        if not positive:
            if i > n - (L - 1):
                # high_boundary_coefs[n-i+1] is the row of node (n-(L-1)) + (n-i) + 1 (same list-order quirk)
                return n - L + 1, capi.fd_weights(d, x[n - (L - 1) + (n - i)], x[n - L:])
            Wf = self._memo(("ufnu", d), lambda: capi.fd_weights_rows(
                d, x[:n - L + 1], np.lib.stride_tricks.sliding_window_view(x, L)))
            return i, Wf[i - 1]
        # REFERENCE QUIRK: the backward table is built for nodes i >= 1+offside but the
        # struct's offside is reset to 0 (upwind_diff_weights.jl:154) and the lookup is stencil_coefs[i - 0]
        # (upwind_difference.jl:157): node i uses the weights computed for node i+offside.
        off = L - 1
        ii = i + off
        if ii > n:
            raise StencilLoweringError("non-uniform backward upwind row past the end of the table (reference BoundsError)")
        Wb = self._memo(("ubnu", d), lambda: capi.fd_weights_rows(
            d, x[off:], np.lib.stride_tricks.sliding_window_view(x, L)))          # row k: node k + 1 + off on taps k + 1 .. k + L
        return i - L + 1, Wb[ii - 1 - off]

    # -- half-offset (CompleteHalfCenteredDifference + get_half_offset_weights_and_stencil) --------
    def half_row(self, d, p, m, periodic, length=None, on_half_grid=False):
        """Row for half point m (between nodes m and m+1).  on_half_grid: the table lives on the grid of
        half points (the outer operator of the nonlinear Laplacian, differential_discretizer.jl:41-53)."""
        x = self.ax.x
        if on_half_grid and not self.ax.uniform:
            x = self._memo("half_nodes", lambda: 0.5 * (self.ax.x[:-1] + self.ax.x[1:]))
        n = len(x) if not self.ax.uniform else (self.ax.n - 1 if on_half_grid else self.ax.n)
        ln = n if length is None else length
        L = p + 2 * (d // 2) + (p % 2)
        bsl, bpc, endpoint = d + p, L // 2, L // 2
        lo_open, up_open = _ends(periodic)
        if self.ax.uniform:
            s = 1.0 / self.ax.dx ** d
            if m <= bpc and not lo_open:
                w = self._memo(("hl", d, p, m), lambda: s * capi.fd_weights(d, 0.5 + m, np.arange(1, bsl + 1, dtype=float)))
                return 1, w
            if m > ln - bpc and not up_open:
                k = ln - m         # high_boundary_coefs[len - m] == reversed low row k (1-based) * (-1)^d
                if k < 1:
                    raise StencilLoweringError("half-offset row requested at the last node")
                w = self._memo(("hh", d, p, k), lambda: (s * capi.fd_weights(d, 0.5 + k, np.arange(1, bsl + 1, dtype=float)))[::-1]
                               * (-1.0) ** d)
                return ln - bsl + 1, w
            w = self._memo(("hi", d, p), lambda: s * capi.fd_weights(d, 0.5, np.arange(1 - endpoint, endpoint + 1, dtype=float)))
            return m + 1 - L // 2, w
        if lo_open or up_open:
            raise StencilLoweringError("periodic boundaries are not supported on non-uniform grids for half-offset stencils")
        hx = self._memo(("half_of", on_half_grid), lambda: 0.5 * (x[:-1] + x[1:]))
        if m <= bpc:
            return 1, capi.fd_weights(d, hx[m - 1], x[:bsl])
        if m > ln - bpc:
            k = ln - m
            if k < 1:
                raise StencilLoweringError("half-offset row requested at the last node")
            return ln - bsl + 1, capi.fd_weights(d, hx[len(hx) - k], x[len(x) - bsl:])
        Wh = self._memo(("hnu", d, p, on_half_grid), lambda: capi.fd_weights_rows(
            d, hx[endpoint - 1:len(x) - endpoint], np.lib.stride_tricks.sliding_window_view(x, 2 * endpoint)))
        return m + 1 - L // 2, Wh[m - endpoint]

    # -- extrapolation pad (BoundaryInterpolatorExtrapolator + central_difference) -------------------
    def extrap_row(self, i):
        p = max(6, self.p)
        n, x = self.ax.n, self.ax.x
        L = p - 1 + p % 2
        bsl, bpc = p, L // 2

        def lag(nodes, k):
            rem = np.delete(nodes, k)
            w = capi.fd_weights(0, nodes[k], rem)
            return np.insert(w, k, 0.0)
        if self.ax.uniform:
            nodes = np.arange(bsl, dtype=float)
            if i <= bpc:
                return 1, lag(nodes, i - 1)
            if i > n - bpc:
                return n - bsl + 1, lag(nodes, n - i)[::-1]
        else:
            dxs = self._memo("dxs", lambda: np.diff(self.ax.x))
            if i <= bpc:
                lx = np.concatenate([[0.0], np.cumsum(dxs[:bsl - 1])])
                return 1, lag(lx, i - 1)
            if i > n - bpc:
                hx = np.cumsum(dxs[n - 1 - bsl:])
                return n - bsl + 1, lag(hx, len(hx) - 1 - (n - i))
        raise StencilLoweringError("extrapolation pad requested away from the boundary")


# ---------------------------------------------------------------------------------------- lowering
class _Tab:
    def __init__(self, tid, L, first, nrows):
        self.id, self.L, self.first, self.nrows = tid, L, first, nrows
        self.rows = {}           # idx -> (start, weights)
        self.core = None         # (lo, hi, off, weights)
        self.score = None        # (lo, hi, off, ntaps): "shape core" -- same taps relative to the node, per-node weights

    def _shape_core(self):
        """Longest contiguous run of rows with the same (first tap - node, number of taps): on a non-uniform grid the
        interior rows of centered_diff_weights.jl:94-103 / upwind_diff_weights.jl:107-133 differ in their weights only."""
        best, run = None, None
        for idx in sorted(self.rows):
            st, w = self.rows[idx]
            key = (st - idx, len(w))
            if run is not None and run[2] == key and run[1] == idx - 1:
                run = (run[0], idx, key)
            else:
                run = (idx, idx, key)
            if best is None or run[1] - run[0] > best[1] - best[0]:
                best = run
        if best is not None:
            self.score = (best[0], best[1], best[2][0], best[2][1])

    def finalize(self, allow_core):
        """Find the contiguous range of rows that share one shifted literal row."""
        if not self.rows:
            return
        self._shape_core()
        if not allow_core:
            return
        groups = {}
        for idx, (st, w) in self.rows.items():
            groups.setdefault((st - idx, tuple(float(v) for v in w)), []).append(idx)
        key, members = max(groups.items(), key=lambda kv: len(kv[1]))
        members.sort()
        # longest contiguous run
        best = (members[0], members[0])
        lo = prev = members[0]
        for k in members[1:]:
            if k != prev + 1:
                lo = k
            prev = k
            if prev - lo > best[1] - best[0]:
                best = (lo, prev)
        if best[1] - best[0] >= 0:
            w = np.zeros(self.L)
            w[:len(key[1])] = key[1]
            self.core = (best[0], best[1], key[0], w)

    def is_core(self, idx):
        return self.core is not None and self.core[0] <= idx <= self.core[1]

    def emit(self, out):
        out.append(f"tab {self.id} {self.L} {self.nrows} {self.first}")
        if self.core is not None:
            lo, hi, off, w = self.core
            out.append(f"core {self.id} {lo} {hi} {off} " + " ".join(_hex(v) for v in w))
        if self.score is not None:
            out.append("score {} {} {} {} {}".format(self.id, *self.score))
        for idx in sorted(self.rows):
            if self.is_core(idx):
                continue
            st, w = self.rows[idx]
            out.append(f"row {self.id} {idx} {st} {len(w)} " + " ".join(_hex(v) for v in w))


class StencilProgram:
    """Result of the lowering: IR text + layout metadata the host layer needs."""

    def __init__(self):
        self.text = ""
        self.nstate = 0
        self.var_names = []
        self.offsets = []
        self.shapes = []          # interior extents per var
        self.ilo = []
        self.ihi = []
        self.axes = []
        self._u0 = None
        self._u0_fn = None        # initial condition is evaluated on first use (1024^3 grids: only when asked for)
        self.tspan = (0.0, 1.0)
        self.params = []
        self.pvals = np.zeros(0)
        self.periodic = []
        self.segments = None      # domains joined by interfaces: per variable {sym, off, n, x}: chart nodes off+1 .. off+n
        self.corebox = None

    @property
    def u0(self):
        if self._u0 is None and self._u0_fn is not None:
            self._u0 = self._u0_fn()
        return self._u0



def _weno_core(rows):
    """Longest run of WENO rows with the centre target on five consecutive nodes (taps i-2 .. i+2), or None."""
    good = sorted(i for i, (s0, T) in rows.items() if T == 3 and s0 == i - 2)
    if not good:
        return None
    runs = []
    a = b = good[0]
    for i in good[1:]:
        if i == b + 1:
            b = i
        else:
            runs.append((a, b)); a = b = i
    runs.append((a, b))
    return max(runs, key=lambda r: r[1] - r[0])


class Lowering:
    def __init__(self, pdesys, disc):
        self.sys, self.disc = pdesys, disc
        self.t = disc.time
        if self.t is None:
            raise StencilLoweringError("steady-state problems (time = nothing) are outside the explicit-RK hot path")
        self.dvs = list(pdesys.dvs)
        self.fns = [d.func for d in self.dvs]
        self.nv = len(self.dvs)
        spatial = [[a for a in d.args if a != self.t] for d in self.dvs]
        self.nd = max(len(sa) for sa in spatial)
        if any(len(sa) not in (0, self.nd) for sa in spatial) or (self.nd != 1 and any(not sa for sa in spatial)):
            raise StencilLoweringError("all dependent variables must have the same number of spatial arguments "
                                       "(variables of t alone are lowered next to 1-D variables)")
        if not 1 <= self.nd <= 3:
            raise StencilLoweringError("1 to 3 spatial dimensions are supported")
        dom = {iv.var: (float(iv.lo), float(iv.hi)) for iv in pdesys.domains}
        self.dom = dom
        self.tspan = dom[self.t]
        self.params = [p for p, _ in pdesys.ps]
        self.pvals = np.array([v for _, v in pdesys.ps], dtype=float)
        if type(disc.grid_align).__name__ not in ("CenterAlignedGrid", "EdgeAlignedGrid"):
            raise StencilLoweringError("center- and edge-aligned grids are lowered (staggered grids are out of scope)")
        self.edge = type(disc.grid_align).__name__ == "EdgeAlignedGrid"
        sch = disc.advection_scheme
        self.weno = type(sch).__name__ == "WENOScheme"
        self.weno_eps = float(getattr(sch, "epsilon", 1e-6))
        self.pu = int(getattr(sch, "order", 1))
        self.eqs, self.bcs = list(pdesys.eqs), list(pdesys.bcs)
        self.segments = None
        # interface neighbours per variable and dimension: [lower, upper] variable index or None
        self.nbr = [[[None, None] for _ in range(self.nd)] for _ in range(self.nv)]
        if any(sa != spatial[0] for sa in spatial):
            self._join_domains(spatial)        # variables on different domains joined by interfaces: one chart axis
        else:
            self.xs = spatial[0]
            self.axes = [Axis(x, dom[x][0], dom[x][1], disc.dxs[x], self.edge) for x in self.xs]
            self.st = [AxisStencils(ax, disc.approx_order, self.pu) for ax in self.axes]
            self.vax = [list(self.axes) for _ in range(self.nv)]
            self.vst = [list(self.st) for _ in range(self.nv)]
            self.voff = [[0] * self.nd for _ in range(self.nv)]
        # BoundaryInterpolatorExtrapolator on a non-uniform grid (extrapolation_weights.jl:84-87): the order-max(6, p)
        # boundary stencil must fit, whether or not a pad ends up using it
        need = max(6, disc.approx_order) + 1
        for v in range(self.nv):
            for ax in self.vax[v]:
                if not ax.uniform and 1 < ax.n < need:
                    raise StencilLoweringError(f"grid has {ax.n} points, but the boundary extrapolation stencil needs at "
                                               f"least {need}. Provide a finer grid for this variable.")
        self.tabs, self.wtabs, self.fn_exprs, self.ghost_lines = [], [], [], []
        self._tabcache = {}
        self._tabsig = {}
        self._classify_bcs()
        # (periodic dimensions need nothing special on an edge-aligned grid: the reference identifies node 1 with node n and
        # wraps taps by n - 1 whatever the alignment -- generate_bc_eqs.jl:35-58, interface_boundary.jl:33-42)
        if self.edge and self.weno:
            raise StencilLoweringError("edge-aligned grids are lowered for centered/upwind schemes")
        self._interiors()

    # -- variables on different domains joined by interfaces (interface_boundary.jl:79-153) -------------------------
    def _join_domains(self, spatial):
        """Two-domain interface boundary conditions `u1(t, b) ~ u2(t, b)` with u1(t, x1), u2(t, x2) on their own
        domains and grids (test/Diffusion/MOL_1D_Linear_Diffusion.jl:887-930, test/Convection_NU/
        MOL_1D_Interface_Upwind_NonUniform.jl:122-170).  The reference wraps a tap that leaves x1's grid through the
        interface onto u2's array (`_wrapinterface`, interface_boundary.jl:79-107) and reads its coordinate in u1's
        chart (`bcoord`, :109-153: the neighbour's grid shifted so that the two edges coincide).  Here the joined
        domains become ONE chart axis -- the grids laid end to end, the shared edge node counted once -- on which
        every variable owns a contiguous node range; the system is rewritten onto the chart coordinate, and a tap
        past the interface is an ordinary ghost rule `u1[node] = 1.0 * u2[node]` (same chart node)."""
        if self.nd != 1:
            raise StencilLoweringError("interfaces between different domains are lowered in one spatial dimension")
        if self.edge:
            raise StencilLoweringError("interfaces between domains are lowered on centre-aligned grids")
        # a variable of t alone (an ODE next to the PDEs, test/Diffusion/MOL_1D_Linear_Diffusion.jl:830-885) lives on a
        # chart segment of one node
        spatial = [sa if sa else [sp.Symbol(f"__point_{self.fns[v]}")] for v, sa in enumerate(spatial)]
        points = {sa[0] for sa in spatial if str(sa[0]).startswith("__point_")}
        syms = []
        for sa in spatial:
            if sa[0] not in syms:
                syms.append(sa[0])
        vsym = [sa[0] for sa in spatial]
        dom, tol = dict(self.dom), 1e-12
        dxs = dict(self.disc.dxs)
        for q in points:
            dom[q], dxs[q] = (0.0, 0.0), np.array([0.0])
        up_link, lo_link = {}, {}                 # symbol -> the symbol joined at its upper / lower end
        rest = []
        for eq in self.bcs:
            L, R = eq.lhs, eq.rhs
            fl, fr = getattr(L, "func", None), getattr(R, "func", None)
            if fl in self.fns and fr in self.fns and fl != fr:
                a, b = self.fns.index(fl), self.fns.index(fr)
                ka = ([k for k, q in enumerate(self.dvs[a].args) if q != self.t] or [None])[0]
                kb = ([k for k, q in enumerate(self.dvs[b].args) if q != self.t] or [None])[0]
                if ka is not None and kb is not None and L.args[ka].is_number and R.args[kb].is_number and vsym[a] != vsym[b]:
                    va, vb = float(L.args[ka]), float(R.args[kb])
                    at_hi = lambda val, x: abs(val - dom[x][1]) <= tol * max(1.0, abs(dom[x][1]))
                    at_lo = lambda val, x: abs(val - dom[x][0]) <= tol * max(1.0, abs(dom[x][0]))
                    if at_hi(va, vsym[a]) and at_lo(vb, vsym[b]):
                        lower_v, upper_v = a, b           # a's upper end meets b's lower end
                    elif at_lo(va, vsym[a]) and at_hi(vb, vsym[b]):
                        lower_v, upper_v = b, a
                    else:
                        raise StencilLoweringError(f"interface {eq} joins two variables at the same end of their domains "
                                                   "(interface_boundary.jl:109-111)")
                    if up_link.setdefault(vsym[lower_v], vsym[upper_v]) != vsym[upper_v] or \
                            lo_link.setdefault(vsym[upper_v], vsym[lower_v]) != vsym[lower_v]:
                        raise StencilLoweringError(f"a domain end is joined to two different domains: {eq}")
                    self.nbr[lower_v][0][1] = upper_v
                    self.nbr[upper_v][0][0] = lower_v
                    continue
            rest.append(eq)
        # chains of joined domains, laid one after the other on the chart axis (domains that are not joined to anything,
        # e.g. u(t, x) and v(t, y) of test/Diffusion/MOL_1D_Linear_Diffusion.jl:693-757, are chains of one)
        heads = [x for x in syms if x not in lo_link]
        order, is_head = [], {}
        for h in heads:
            order.append(h); is_head[h] = True
            while order[-1] in up_link:
                nxt = up_link[order[-1]]
                if nxt in order:
                    raise StencilLoweringError("the domains joined by interfaces must form open chains (no rings)")
                order.append(nxt); is_head[nxt] = False
        if len(order) != len(syms):
            raise StencilLoweringError("the domains joined by interfaces must form open chains (no rings)")
        X = order[0]
        # _check_interface_boundarymap (MOL_discretization.jl:55-95)
        isvec = {x: np.ndim(dxs[x]) > 0 for x in order}
        for a, b in up_link.items():
            if not self.weno and self.pu > 1 and (isvec[a] or isvec[b]):
                raise StencilLoweringError(f"UpwindScheme(order={self.pu}) is not supported with interface boundary "
                                           "conditions on nonuniform grids")
            if isvec[a] != isvec[b]:
                raise StencilLoweringError(f"the interface between {a} and {b} mixes a scalar step size with a nonuniform "
                                           "grid vector")
            if not isvec[a] and dxs[a] != dxs[b]:
                raise StencilLoweringError(f"the step size of the connected variables {a} and {b} must be the same")
        shift, off, segax = {}, {}, {}
        xs_chart = []
        for x in order:
            lo, hi = dom[x]
            own = Axis(x, lo, hi, dxs[x], False)
            if is_head[x]:
                shift[x] = 0.0
            else:
                # bcoord (interface_boundary.jl:109-153): the neighbour's grid moved so that the two edge nodes coincide
                shift[x] = xs_chart[-1] - own.x[0]
                scale = max(abs(own.x[-1] - own.x[0]), abs(xs_chart[-1] - xs_chart[0]))
                if abs(shift[x]) <= 1e-12 * scale:
                    shift[x] = 0.0
                elif isvec[x] and not self.weno and abs(shift[x]) > math.sqrt(np.finfo(float).eps) * scale:
                    raise StencilLoweringError(f"the physical coordinates at the interface of {x} must match for "
                                               "nonuniform grids (MOL_discretization.jl:79-86)")
            spec = dxs[x]
            if isvec[x]:
                spec = np.asarray(spec, dtype=float) + shift[x]
            ax = Axis(X, lo + shift[x], hi + shift[x], spec, False)
            segax[x] = ax
            off[x] = len(xs_chart) if is_head[x] else len(xs_chart) - 1
            xs_chart += list(ax.x if is_head[x] else ax.x[1:])
        chart = Axis(X, xs_chart[0], xs_chart[-1], np.array(xs_chart), False)    # a lookup table of coordinates: rows are per variable
        self.xs, self.axes = [X], [chart]
        self.st = [AxisStencils(chart, self.disc.approx_order, self.pu)]
        self.vax = [[segax[vsym[v]]] for v in range(self.nv)]
        stn = {x: AxisStencils(segax[x], self.disc.approx_order, self.pu) for x in order}
        self.vst = [[stn[vsym[v]]] for v in range(self.nv)]
        self.voff = [[off[vsym[v]]] for v in range(self.nv)]
        self.segments = [dict(sym=vsym[v], off=off[vsym[v]], n=segax[vsym[v]].n, x=segax[vsym[v]].x - shift[vsym[v]])
                         for v in range(self.nv)]
        # rewrite the system onto the chart coordinate: u_k(t, x_k) -> u_k(t, X), bare x_k -> X - shift_k
        tpos = [[k for k, q in enumerate(d.args) if q == self.t] for d in self.dvs]

        def canon(e):
            f = getattr(e, "func", None)
            if f in self.fns:
                v = self.fns.index(f)
                args = []
                for k, q in enumerate(e.args):
                    if k in tpos[v]:
                        args.append(q)
                    elif q == vsym[v]:
                        args.append(X)
                    elif q.is_number:
                        args.append(q + shift[vsym[v]] if shift[vsym[v]] != 0.0 else q)
                    else:
                        raise StencilLoweringError(f"spatial argument of {e} is neither its variable nor a number")
                return f(*args)
            if isinstance(e, sp.Derivative):
                return sp.Derivative(canon(e.expr), *[(X if var in syms else var, cnt) for var, cnt in e.variable_count])
            if e in syms:
                return X - shift[e] if shift[e] != 0.0 else X
            if not e.args:
                return e
            return e.func(*[canon(q) for q in e.args])
        Eqn = type(self.eqs[0])
        self.eqs = [Eqn(canon(eq.lhs), canon(eq.rhs)) for eq in self.eqs]
        self.bcs = [Eqn(canon(eq.lhs), canon(eq.rhs)) for eq in rest]
        self.dvs = [canon(d) for d in self.dvs]

    # -- boundaries (PDEBase.parse_bcs analogue) -------------------------------------------------
    def _calls(self, expr, fn):
        return [a for a in expr.atoms(sp.core.function.AppliedUndef) if a.func == fn]

    def _classify_bcs(self):
        nv, nd = self.nv, self.nd
        self.per = [[False] * nd for _ in range(nv)]
        self.bc = [[[None, None] for _ in range(nd)] for _ in range(nv)]
        self.bcx = {}            # (v, j, side) -> all conditions at that end when there is more than one
        self.ic = [None] * nv
        for eq in self.bcs:
            done = False
            for v, (dv, fn) in enumerate(zip(self.dvs, self.fns)):
                for call in self._calls(eq.lhs, fn) + self._calls(eq.rhs, fn):
                    fixed = [(k, a) for k, a in enumerate(call.args) if a.is_number]
                    if not fixed:
                        continue
                    k, val = fixed[0]
                    canon = dv.args[k]
                    if canon == self.t:
                        if eq.lhs != call:
                            raise StencilLoweringError(f"initial condition must read u(t0, ...) ~ expr: {eq}")
                        self.ic[v] = eq.rhs
                    else:
                        j = self.xs.index(canon)
                        lo, hi = self.vax[v][j].lo, self.vax[v][j].hi
                        both = (getattr(eq.lhs, "func", None) == fn and getattr(eq.rhs, "func", None) == fn)
                        if both and {float(eq.lhs.args[k]), float(eq.rhs.args[k])} == {lo, hi}:
                            self.per[v][j] = True
                        else:
                            upper = abs(float(val) - hi) <= 1e-12 * max(1.0, abs(hi))
                            if not upper and abs(float(val) - lo) > 1e-12 * max(1.0, abs(lo)):
                                raise StencilLoweringError(f"boundary condition not on a domain boundary: {eq}")
                            if self.nbr[v][j][int(upper)] is not None:
                                raise StencilLoweringError(f"boundary condition at an interface end: {eq}")
                            if self.bc[v][j][int(upper)] is None:
                                self.bc[v][j][int(upper)] = eq
                            else:       # several conditions at one end (higher-order PDEs: test/Higher_Order)
                                self.bcx.setdefault((v, j, int(upper)), [self.bc[v][j][int(upper)]]).append(eq)
                    done = True
                    break
                if done:
                    break
            if not done:
                raise StencilLoweringError(f"could not classify boundary condition {eq}")
        # open ends: a periodic or interface boundary sits there, so no one-sided boundary row is ever selected
        self.open = [[(self.per[v][j] or self.nbr[v][j][0] is not None, self.per[v][j] or self.nbr[v][j][1] is not None)
                      for j in range(nd)] for v in range(nv)]

    def _eq_var(self, eq):
        for D in (eq.lhs - eq.rhs).atoms(sp.Derivative):
            if D.variables == (self.t,) and D.expr in self.dvs:
                return self.dvs.index(D.expr)
        raise StencilLoweringError(f"equation has no time derivative of a dependent variable: {eq}")

    def _interiors(self):
        """interior_map.jl:1-10,89-115,117-139."""
        self.eq_of = {}
        for eq in self.eqs:
            v = self._eq_var(eq)
            if v in self.eq_of:
                raise StencilLoweringError("two equations for the same variable")
            self.eq_of[v] = eq
        if len(self.eq_of) != self.nv:
            raise StencilLoweringError("need one evolution equation per dependent variable")
        self.ilo, self.ihi, self.vlo, self.vup, self.ext = [], [], [], [], []
        for v in range(self.nv):
            resid = self.eq_of[v].lhs - self.eq_of[v].rhs
            lo, up, le, ue = [], [], [], []
            for j, x in enumerate(self.xs):
                if self.per[v][j]:
                    l, u_ = 1, 0
                else:   # clip_interior!! (interior_map.jl:1-10): an interface clips its lower end only
                    # every condition at an end clips one more node there
                    l = len(self.bcx.get((v, j, 0), [0] * int(self.bc[v][j][0] is not None))) + int(self.nbr[v][j][0] is not None)
                    u_ = len(self.bcx.get((v, j, 1), [0] * int(self.bc[v][j][1] is not None)))
                e = 0
                for Dn in resid.atoms(sp.Derivative):
                    for var, cnt in Dn.variable_count:
                        if var == x and int(cnt) == 1 and self.weno and self.vax[v][j].uniform:
                            e = 2
                # calculate_stencil_extents (interior_map.jl:117-139): an open end has no extent
                lo.append(l); up.append(u_)
                le.append(0 if self.open[v][j][0] else e); ue.append(0 if self.open[v][j][1] else e)
            self.vlo.append(lo); self.vup.append(up); self.ext.append((le, ue))
            low = [max(a, b) for a, b in zip(lo, le)]
            upp = [max(a, b) for a, b in zip(up, ue)]
            for j in range(self.nd):
                if low[j] + upp[j] + 1 > self.vax[v][j].n:
                    raise StencilLoweringError("The domain is too small to support the requested discretization")
            self.ilo.append([self.voff[v][j] + 1 + low[j] for j in range(self.nd)])
            self.ihi.append([self.voff[v][j] + self.vax[v][j].n - upp[j] for j in range(self.nd)])

    # -- tables ----------------------------------------------------------------------------------------
    def _new_tab(self, key, L, first, nrows, rowfn, allow_core):
        if key in self._tabcache:
            return self._tabcache[key]
        T = _Tab(len(self.tabs), L, first, nrows)
        for idx in range(first, first + nrows):
            st, w = rowfn(idx)
            if len(w) > L:
                raise StencilLoweringError("stencil row longer than its table")
            T.rows[idx] = (int(st), np.asarray(w, dtype=float))
        # identical tables are shared (e.g. the u and v operators of a system with the same boundary types): the fused
        # kernel then loads each table weight once per node for all equations
        sig = (L, first, nrows, tuple((idx, T.rows[idx][0], T.rows[idx][1].tobytes()) for idx in sorted(T.rows)))
        dup = self._tabsig.get(sig)
        if dup is not None:
            self._tabcache[key] = dup
            return dup
        T.finalize(allow_core)
        self.tabs.append(T)
        self._tabcache[key] = T
        self._tabsig[sig] = T
        return T

    def _local(self, u, j, ev):
        """Stencil rows are built on the variable's own grid (local node numbers 1..n) and stored with chart node
        numbers: chart node = local node + offset (0 unless domains are joined by interfaces)."""
        if self.vax[ev][j] is not self.vax[u][j]:
            raise StencilLoweringError("a derivative of a variable that lives on another domain than its equation")
        return self.vst[u][j], self.open[u][j], self.voff[u][j]

    def _chart_coord(self, u, j):
        """bcoord (interface_boundary.jl:109-153): coordinate of the raw local node r of variable u in u's own chart --
        across an interface the neighbour's grid (same chart axis), across a periodic seam the grid shifted by the
        period."""
        ax, off, chart = self.vax[u][j], self.voff[u][j], self.axes[j]

        def coord(r):
            if self.per[u][j]:
                Lp = ax.x[-1] - ax.x[0]
                if r < 1:
                    return ax.x[r + ax.n - 2] - Lp
                if r > ax.n:
                    return ax.x[r - ax.n] + Lp
                return ax.x[r - 1]
            if not 1 <= r + off <= chart.n:
                raise StencilLoweringError("stencil tap past the end of the joined domains")
            return chart.x[r + off - 1]
        return coord

    def tab_centered(self, u, j, d, ev):
        st, opn, off = self._local(u, j, ev)
        p = self.disc.approx_order
        L = max(d + p - 1 + (d + p) % 2, d + p)
        lo, hi = self.ilo[ev][j], self.ihi[ev][j]
        if off or any(w_ is not None for w_ in self.nbr[u][j]):
            self._check_interface_spacing(u, j, d)

        def row(i):
            s0, w = st.centered_row(d, i - off, opn)
            return s0 + off, w
        return self._new_tab(("c", u, j, d, lo, hi), L, lo, hi - lo + 1, row, self.vax[u][j].uniform)

    def _check_interface_spacing(self, u, j, d):
        """validate_interface_orders (interior_map.jl:33-52): only first-order derivatives are coordinate-aware across
        an interface; higher orders need identical step sizes on both sides."""
        for w_ in self.nbr[u][j]:
            if w_ is None:
                continue
            a, b = self.vax[u][j], self.vax[w_][j]
            if d > 1 and not (a.uniform and b.uniform and abs(a.dx - b.dx) <= 1e-12 * abs(a.dx)):
                raise StencilLoweringError("derivative orders > 1 across an interface need identical uniform step sizes "
                                           "on both sides (interior_map.jl:33-52)")

    def tab_upwind(self, u, j, d, ev, positive):
        st, opn, off = self._local(u, j, ev)
        lo, hi = self.ilo[ev][j], self.ihi[ev][j]
        coord = self._chart_coord(u, j)
        if off or any(w_ is not None for w_ in self.nbr[u][j]):
            self._check_interface_spacing(u, j, d)

        def row(i):
            s0, w = st.upwind_row(d, i - off, positive, opn, coord)
            return s0 + off, w
        return self._new_tab(("w", u, j, d, positive, lo, hi), d + self.pu, lo, hi - lo + 1, row, self.vax[u][j].uniform)

    def wtab(self, u, j, ev):
        st, (lo_open, up_open), off = self._local(u, j, ev)
        n = self.vax[u][j].n
        lo, hi = self.ilo[ev][j], self.ihi[ev][j]
        rows = {}
        for ic in range(lo, hi + 1):
            i = ic - off
            if i <= 2 and not lo_open:
                rows[ic] = (1 + off, i)
            elif i > n - 2 and not up_open:
                rows[ic] = (n - 4 + off, 5 - (n - i))
            else:
                if (lo_open or up_open) and n - 1 < 5:
                    raise StencilLoweringError("WENO needs at least 6 grid points to wrap across a periodic boundary")
                rows[ic] = (ic - 2, 3)
        if self.vax[u][j].uniform and any(T != 3 for _, T in rows.values()):
            raise StencilLoweringError("uniform WENO is only defined on the interior (extent 2)")
        wid = len(self.wtabs)
        self.wtabs.append((wid, lo, hi, rows))
        return wid

    # -- expression -> RPN ---------------------------------------------------------------------------
    def rpn(self, e, ops=None, allow_fields=True):
        ops = ops or {}
        out = []

        def go(e):
            if e in ops:
                out.append(ops[e]); return
            if e.is_Number or isinstance(e, sp.NumberSymbol):
                out.append("c:" + _hex(float(e))); return
            if e == self.t:
                out.append("t"); return
            if e in self.xs:
                out.append(f"x:{self.xs.index(e)}"); return
            if e in self.params:
                out.append(f"p:{self.params.index(e)}"); return
            if e in self.dvs:
                if not allow_fields:
                    raise StencilLoweringError(f"dependent variable inside boundary data: {e}")
                w_ = self.dvs.index(e)
                out.append(f"s:{w_}" if w_ in getattr(self, "_point_reads", ()) else f"u:{w_}"); return
            if e is sp.true or e is sp.false:
                out.append("c:" + _hex(1.0 if e is sp.true else 0.0)); return
            if isinstance(e, sp.Add):
                go(e.args[0])
                for a in e.args[1:]:
                    go(a); out.append("+")
                return
            if isinstance(e, sp.Mul):
                c, rest = e.as_coeff_Mul()
                if c == -1 and rest != 1:
                    go(rest); out.append("neg"); return
                go(e.args[0])
                for a in e.args[1:]:
                    go(a); out.append("*")
                return
            if isinstance(e, sp.Pow):
                b, x = e.args
                if x.is_Integer:
                    go(b); out.append(f"powi:{int(x)}"); return
                if x == sp.Rational(1, 2):
                    go(b); out.append("sqrt"); return
                go(b); go(x); out.append("pow"); return
            if isinstance(e, sp.Piecewise):
                def pw(args):
                    val, cond = args[0]
                    if cond is sp.true or len(args) == 1:
                        go(val); return
                    go(cond); go(val); pw(args[1:]); out.append("sel")
                pw(list(e.args)); return
            if isinstance(e, (sp.And, sp.Or)):
                go(e.args[0])
                for a in e.args[1:]:
                    go(a); out.append("and" if isinstance(e, sp.And) else "or")
                return
            if isinstance(e, sp.Not):
                go(e.args[0]); out.append("not"); return
            if isinstance(e, sp.core.relational.Relational):
                name = {sp.Gt: "gt", sp.Ge: "ge", sp.Lt: "lt", sp.Le: "le", sp.Eq: "eq", sp.Ne: "ne"}[type(e)]
                go(e.lhs); go(e.rhs); out.append(name); return
            if isinstance(e, (sp.Max, sp.Min)):
                go(e.args[0])
                for a in e.args[1:]:
                    go(a); out.append("max" if isinstance(e, sp.Max) else "min")
                return
            fname = {sp.exp: "exp", sp.log: "log", sp.sin: "sin", sp.cos: "cos", sp.tan: "tan", sp.sinh: "sinh",
                     sp.cosh: "cosh", sp.tanh: "tanh", sp.Abs: "abs", sp.sign: "sign", sp.asin: "asin",
                     sp.acos: "acos", sp.atan: "atan", sp.erf: "erf"}.get(type(e))
            if fname:
                go(e.args[0]); out.append(fname); return
            raise StencilLoweringError(f"expression node not supported by the stencil program: {type(e).__name__}: {e}")
        go(sp.sympify(e))
        return out

    # -- term lowering (generate_finite_difference_rules.jl precedence) ---------------------------------
    @staticmethod
    def split_additive(expr):
        out = []
        for term in sp.Add.make_args(expr):
            c, rest = term.as_coeff_Mul()
            if isinstance(rest, sp.Add):
                out += [c * q for q in Lowering.split_additive(rest)]
            else:
                out.append(term)
        return out

    def _op(self, ops, token):
        s = sp.Symbol(f"__op{len(ops)}")
        ops[s] = token
        return s

    def _L(self, ops, T, u, j):
        return self._op(ops, f"L:{T.id}:{u}:{j}")

    def _nonlinlap(self, ops, inner, u, j, ev):
        """Returns the placeholder for Dx(inner * Dx(u)) at the nodes of equation ev."""
        p = self.disc.approx_order
        if self.segments is not None:
            raise StencilLoweringError("the nonlinear Laplacian is not lowered across interfaces between domains")
        st, per, n = self.st[j], self.per[u][j], self.axes[j].n
        lo, hi = self.ilo[ev][j], self.ihi[ev][j]
        # outer operator: half-offset first derivative on the clipped grid, evaluated at II - 1
        L_o = max(p + (p % 2), 1 + p)
        outer = self._new_tab(("no", u, j, lo, hi), L_o, lo, hi - lo + 1,
                              lambda i: st.half_row(1, p, i - 1, per, length=n - 1, on_half_grid=True),
                              self.axes[j].uniform)
        ms = sorted({m for (s0, w) in outer.rows.values() for m in range(s0, s0 + len(w))})
        mlo, mhi = ms[0], ms[-1]
        pi = max(4, p)
        L_i = max(pi + (pi % 2), pi)
        interp = self._new_tab(("ni", j, per, mlo, mhi), L_i, mlo, mhi - mlo + 1,
                               lambda m: st.half_row(0, pi, m, per), self.axes[j].uniform)
        L_d = max(p + (p % 2), 1 + p)
        deriv = self._new_tab(("nd", j, per, mlo, mhi), L_d, mlo, mhi - mlo + 1,
                              lambda m: st.half_row(1, p, m, per), self.axes[j].uniform)
        if inner.atoms(sp.Derivative):
            raise StencilLoweringError("derivatives inside the nonlinear-Laplacian coefficient are not lowered")
        fid = len(self.fn_exprs)
        self.fn_exprs.append(self.rpn(inner))
        return self._op(ops, f"N:{u}:{j}:{fid}:{interp.id}:{deriv.id}:{outer.id}")

    def _lower_generic(self, expr, ops, ev):
        subs = {}
        for Dn in expr.atoms(sp.Derivative):
            if Dn.expr in self.dvs and len(Dn.variable_count) == 2 and all(int(cn) == 1 and var in self.xs
                                                                           for var, cn in Dn.variable_count):
                # mixed derivative Dx Dy u: product of the centred first-derivative rows of the two dimensions
                # (generate_mixed_rules / mixed_central_difference, 2nd_order_mixed_deriv.jl:5-55)
                u = self.dvs.index(Dn.expr)
                jx, jy = (self.xs.index(var) for var, _ in Dn.variable_count)
                if self.segments is not None or self.edge:
                    raise StencilLoweringError("mixed derivatives are lowered on centre-aligned grids of one domain")
                subs[Dn] = self._op(ops, f"M:{self.tab_centered(u, jx, 1, ev).id}:{self.tab_centered(u, jy, 1, ev).id}:{u}:{jx}:{jy}")
                continue
            if Dn.expr not in self.dvs or len(Dn.variable_count) != 1:
                raise StencilLoweringError(f"derivative pattern not supported: {Dn}")
            x, d = Dn.variable_count[0]
            d = int(d)
            if x not in self.xs:
                raise StencilLoweringError(f"derivative with respect to {x}")
            j, u = self.xs.index(x), self.dvs.index(Dn.expr)
            if d % 2 == 0:
                subs[Dn] = self._L(ops, self.tab_centered(u, j, d, ev), u, j)
            elif self.weno and d == 1:
                wid = self.wtab(u, j, ev)
                dx = self.vax[u][j].dx if self.vax[u][j].uniform else 0.0
                subs[Dn] = self._op(ops, f"W:{wid}:{u}:{j}:{_hex(self.weno_eps)}:{_hex(dx)}")
            else:
                subs[Dn] = self._L(ops, self.tab_upwind(u, j, d, ev, True), u, j)
        return expr.xreplace(subs)

    def _lower_term(self, term, ops, ev):
        factors = list(sp.Mul.make_args(term))
        for k, f in enumerate(factors):
            if isinstance(f, sp.Derivative) and len(f.variable_count) == 1 and int(f.variable_count[0][1]) == 1:
                r = f.variable_count[0][0]
                inner = list(sp.Mul.make_args(f.expr))
                dus = [q for q in inner if isinstance(q, sp.Derivative) and q.expr in self.dvs
                       and q.variable_count == ((r, 1),)]
                if len(dus) == 1 and r in self.xs:
                    j, u = self.xs.index(r), self.dvs.index(dus[0].expr)
                    rest_in = [q for q in inner if q is not dus[0]]
                    others = factors[:k] + factors[k + 1:]
                    if sp.Pow(r, -2) in others and sp.Pow(r, 2) in rest_in:
                        # spherical_diffusion (spherical_laplacian.jl:10-42)
                        others.remove(sp.Pow(r, -2)); rest_in.remove(sp.Pow(r, 2))
                        a = sp.Mul(*rest_in)
                        d2 = self._L(ops, self.tab_centered(u, j, 2, ev), u, j)
                        d1 = self._L(ops, self.tab_centered(u, j, 1, ev), u, j)
                        nl = self._nonlinlap(ops, a, u, j, ev)
                        sph = sp.Piecewise((6 * a * d2, sp.Abs(r) <= 1e-6), (a * (d1 / r + nl), True))
                        return sp.Mul(*others) * sph
                    nl = self._nonlinlap(ops, sp.Mul(*rest_in), u, j, ev)
                    return self._lower_generic(sp.Mul(*others), ops, ev) * nl
        for k, f in enumerate(factors):
            if isinstance(f, sp.Derivative) and f.expr in self.dvs and len(f.variable_count) == 1:
                x, d = f.variable_count[0]
                d = int(d)
                if x in self.xs and d % 2 == 1 and not (self.weno and d == 1) and len(factors) > 1:
                    j, u = self.xs.index(x), self.dvs.index(f.expr)
                    coef = sp.Mul(*(factors[:k] + factors[k + 1:]))
                    if coef.atoms(sp.Derivative):
                        raise StencilLoweringError("derivatives inside an upwind coefficient are not supported "
                                                   "(upwind_difference.jl:188)")
                    bwd = self._L(ops, self.tab_upwind(u, j, d, ev, True), u, j)
                    fwd = self._L(ops, self.tab_upwind(u, j, d, ev, False), u, j)
                    return sp.Piecewise((coef * bwd, coef > 0), (coef * fwd, True))
        return self._lower_generic(term, ops, ev)

    # -- ghost rules -------------------------------------------------------------------------------------
    def _ghosts(self):
        """value(node) = G(t, x) + sum_k a_k u[tap_k]; rules keyed (v, j, node)."""
        rules = {}
        for v in range(self.nv):
            for j, x in enumerate(self.xs):
                if self.per[v][j]:
                    continue
                n, off = self.vax[v][j].n, self.voff[v][j]
                for side in (0, 1):
                    eq = self.bc[v][j][side]
                    if eq is None:
                        continue
                    node = off + (n if side else 1)
                    eqs = self.bcx.get((v, j, side), [eq])
                    nodes = [node - k if side else node + k for k in range(len(eqs))]
                    try:
                        if len(eqs) > 1:
                            rules.update({(v, j, nd_): r for nd_, r in self._solve_bc_set(eqs, v, j, nodes).items()})
                        else:
                            rules[(v, j, node)] = self._solve_bc(eq, v, j, node)
                    except StencilLoweringError as err:
                        # A derivative condition next to extrapolation pads (uniform WENO5 with a Neumann / Robin end): its
                        # one-sided row reads the pad node, whose extrapolation row reads the edge node -- the reference
                        # hands both algebraic equations to ModelingToolkit; here they are solved together.
                        pads = self._pad_nodes(v, j, bool(side))
                        if "boundary node" not in str(err) or not pads:
                            raise
                        rules.update({(v, j, nd_): r for nd_, r in self._solve_bc_set(eqs, v, j, nodes, pads).items()})
        # interfaces (generate_bc_eqs.jl:35-58, interface_boundary.jl:79-107): past its interface end a variable reads
        # its neighbour at the same chart node; the lower variable owns the shared edge node
        REACH = 8
        for v in range(self.nv):
            for j in range(self.nd):
                off, n = self.voff[v][j], self.vax[v][j].n
                wl, wu = self.nbr[v][j]
                if wl is not None:
                    for node in range(off + 1, max(off + 1 - REACH, self.voff[wl][j]), -1):
                        rules[(v, j, node)] = (sp.Integer(0), {(wl, node): 1.0})
                if wu is not None:
                    for node in range(off + n + 1, min(off + n + REACH, self.voff[wu][j] + self.vax[wu][j].n) + 1):
                        rules[(v, j, node)] = (sp.Integer(0), {(wu, node): 1.0})
        # extrapolation pads (generate_extrap_eqs!, generate_bc_eqs.jl:336-392)
        for v in range(self.nv):
            le, ue = self.ext[v]
            for j in range(self.nd):
                if self.per[v][j]:
                    continue
                n, off = self.vax[v][j].n, self.voff[v][j]
                for upper, e, vl in ((False, le[j], self.vlo[v][j]), (True, ue[j], self.vup[v][j])):
                    ninterp = e - vl
                    while ninterp >= vl:
                        node = off + ((n - ninterp) if upper else (1 + ninterp))
                        ninterp -= 1
                        if self.ilo[v][j] <= node <= self.ihi[v][j] or (v, j, node) in rules:
                            continue
                        if vl == 0:
                            raise StencilLoweringError("extrapolation pad next to an unconstrained boundary node")
                        st, w = self.vst[v][j].extrap_row(node - off)
                        st += off
                        G, taps = sp.Integer(0), {}
                        for k, wk in enumerate(w):
                            tp = st + k
                            if wk == 0.0:
                                continue
                            if self.ilo[v][j] <= tp <= self.ihi[v][j]:
                                taps[(v, tp)] = taps.get((v, tp), 0.0) + float(wk)
                            else:
                                if (v, j, tp) not in rules:
                                    raise StencilLoweringError("extrapolation pad taps an undefined boundary node")
                                G2, t2 = rules[(v, j, tp)]
                                G = G + float(wk) * G2
                                for key, a in t2.items():
                                    taps[key] = taps.get(key, 0.0) + float(wk) * a
                        rules[(v, j, node)] = (G, taps)
        return rules

    def _solve_bc(self, eq, v, j, node):
        """Solve the (affine) boundary equation for the edge node (generate_bc_eqs.jl:313-328).  Centre-aligned grid
        (boundary_value_maps, :238-311): u(t, x_b) is the edge node itself and Dx^d u(t, x_b) the one-sided row of the
        centred operator there.  Edge-aligned grid (:79-161): the boundary lies half-way between the first / last two
        nodes, u(t, x_b) becomes the interpolation row (CompleteHalfCenteredDifference(0, max(4, p))) and Dx^d u(t, x_b)
        the half-offset derivative row at that half point (index 1 or len - 1: newindex(...; shift = true))."""
        x, ax = self.xs[j], self.axes[j]               # chart axis: ax.x[node - 1] is the boundary coordinate
        off, n = self.voff[v][j], self.vax[v][j].n      # rows on the variable's own grid, local node = node - off
        stv = self.vst[v][j]
        half = 1 if node - off == 1 else n - 1
        resid = eq.lhs - eq.rhs
        Ub = sp.Symbol("__Ub")
        tapsyms = {}

        def U(w_, tp):
            if w_ == v and tp == node:
                return Ub
            if not (self.ilo[w_][j] <= tp <= self.ihi[w_][j]):
                raise StencilLoweringError(f"boundary condition {eq} couples to another boundary node")
            return tapsyms.setdefault((w_, tp), sp.Symbol(f"__U_{w_}_{tp}"))
        subs = {}
        for Dn in resid.atoms(sp.Derivative):
            call = Dn.expr
            if getattr(call, "func", None) not in self.fns or len(Dn.variable_count) != 1 or Dn.variable_count[0][0] != x:
                raise StencilLoweringError(f"boundary derivative not supported: {Dn}")
            w_ = self.fns.index(call.func)
            d = int(Dn.variable_count[0][1])
            if self.edge:
                st, w = stv.half_row(d, self.disc.approx_order, half, False)
            else:
                st, w = stv.centered_row(d, node - off, False)
            st += off
            subs[Dn] = sum(float(wk) * U(w_, st + k) for k, wk in enumerate(w))
        resid = resid.xreplace(subs)
        for w_, fn in enumerate(self.fns):
            for call in self._calls(resid, fn):
                if self.edge:
                    st, w = stv.half_row(0, max(4, self.disc.approx_order), half, False)
                    resid = resid.xreplace({call: sum(float(wk) * U(w_, st + k) for k, wk in enumerate(w))})
                else:
                    resid = resid.xreplace({call: U(w_, node)})
        resid = resid.xreplace({x: sp.Float((ax.hi if node == n else ax.lo) if self.edge else ax.x[node - 1])})   # (edge: off = 0)
        resid = sp.expand(resid)
        A = sp.diff(resid, Ub)
        if A == 0 or A.has(Ub) or any(A.has(s) for s in tapsyms.values()):
            raise StencilLoweringError(f"boundary condition is not affine in the boundary value: {eq}")
        # Coefficients may depend on parameters, t and the coordinates along the boundary (a Robin coefficient that is
        # a parameter, a time-dependent mixing ratio ...): they stay symbolic and are emitted as expressions (`ghostx`).
        # They may not depend on field values.
        fieldsyms = set(tapsyms.values()) | {Ub}
        taps, rest = {}, resid - A * Ub
        if A.free_symbols & fieldsyms or any(A.has(fn) for fn in self.fns):
            raise StencilLoweringError(f"boundary condition is not affine in the boundary value: {eq}")
        for key, s in tapsyms.items():
            ck = sp.diff(rest, s)
            if ck.free_symbols & fieldsyms:
                raise StencilLoweringError(f"boundary condition is not affine: {eq}")
            rest = rest - ck * s
            a = sp.simplify(-ck / A) if (ck.free_symbols or A.free_symbols) else -ck / A
            if a.free_symbols:
                taps[key] = a
            elif float(a) != 0.0:
                taps[key] = float(a)
        rest = sp.expand(rest)
        if rest.has(Ub) or any(rest.has(s) for s in tapsyms.values()):
            raise StencilLoweringError(f"boundary condition is not affine: {eq}")
        return (-rest / A, taps)

    def _pad_nodes(self, v, j, upper):
        """Chart nodes of the extrapolation pads at one end (generate_extrap_eqs!, generate_bc_eqs.jl:336-392)."""
        le, ue = self.ext[v]
        n, off = self.vax[v][j].n, self.voff[v][j]
        e, vl = (ue[j], self.vup[v][j]) if upper else (le[j], self.vlo[v][j])
        out, ninterp = [], e - vl
        while ninterp >= vl and vl > 0:
            node = off + ((n - ninterp) if upper else (1 + ninterp))
            ninterp -= 1
            if not (self.ilo[v][j] <= node <= self.ihi[v][j]):
                out.append(node)
        return out

    def _solve_bc_set(self, eqs, v, j, nodes, pads=()):
        """Several boundary conditions at one end (u, Dx u, Dxx u of a third-order PDE ...): the reference clips one node
        per condition (interior_map.jl:1-10), writes EVERY condition at the edge node with the one-sided rows there
        (boundary_value_maps, generate_bc_eqs.jl:238-311) and lets the m clipped nodes be the m unknowns of that affine
        system.  Solved here once, in double precision: node_k = G_k(t, x) + sum_taps a_k,tap u[tap]."""
        if self.edge:
            raise StencilLoweringError("several boundary conditions at one end are lowered on centre-aligned grids")
        x, ax = self.xs[j], self.axes[j]
        off, stv = self.voff[v][j], self.vst[v][j]
        edge = nodes[0]
        nodes = list(nodes) + list(pads)               # pads: extrapolation-pad nodes solved together with the conditions
        Ub = {nd_: sp.Symbol(f"__Ub{nd_}") for nd_ in nodes}
        tapsyms = {}

        def U(w_, tp):
            if w_ == v and tp in Ub:
                return Ub[tp]
            if not (self.ilo[w_][j] <= tp <= self.ihi[w_][j]):
                raise StencilLoweringError(f"boundary conditions at node {edge} couple to another boundary node")
            return tapsyms.setdefault((w_, tp), sp.Symbol(f"__U_{w_}_{tp}"))
        resids = []
        for eq in eqs:
            resid = eq.lhs - eq.rhs
            subs = {}
            for Dn in resid.atoms(sp.Derivative):
                call = Dn.expr
                if getattr(call, "func", None) not in self.fns or len(Dn.variable_count) != 1 or Dn.variable_count[0][0] != x:
                    raise StencilLoweringError(f"boundary derivative not supported: {Dn}")
                st, w = stv.centered_row(int(Dn.variable_count[0][1]), edge - off, False)
                subs[Dn] = sum(float(wk) * U(self.fns.index(call.func), st + off + k) for k, wk in enumerate(w))
            resid = resid.xreplace(subs)
            for w_, fn in enumerate(self.fns):
                for call in self._calls(resid, fn):
                    resid = resid.xreplace({call: U(w_, edge)})
            resids.append(sp.expand(resid.xreplace({x: sp.Float(ax.x[edge - 1])})))
        for pad in pads:                               # u[pad] = sum_k w_k u[tap_k]  (weight 0 at the pad itself)
            st, w = stv.extrap_row(pad - off)
            resids.append(sp.expand(Ub[pad] - sum(float(wk) * U(v, st + off + k) for k, wk in enumerate(w) if wk != 0.0)))
        fieldsyms = set(tapsyms.values()) | set(Ub.values())
        A = np.zeros((len(resids), len(nodes)))
        rest = []
        for k, r in enumerate(resids):
            for i, nd_ in enumerate(nodes):
                a = sp.diff(r, Ub[nd_])
                if a.free_symbols:
                    raise StencilLoweringError("several boundary conditions at one end need constant coefficients")
                A[k, i] = float(a)
                r = r - a * Ub[nd_]
            r = sp.expand(r)
            if r.free_symbols & set(Ub.values()):
                raise StencilLoweringError("boundary condition is not affine in the boundary values")
            rest.append(r)
        if abs(np.linalg.det(A)) < 1e-300:
            raise StencilLoweringError("the boundary conditions at one end do not determine the clipped nodes")
        Ainv = np.linalg.inv(A)
        out = {}
        for i, nd_ in enumerate(nodes):
            val = sp.expand(sum(-float(Ainv[i, k]) * rest[k] for k in range(len(resids))))
            taps = {}
            for key, sym in tapsyms.items():
                a = sp.diff(val, sym)
                if a.free_symbols & fieldsyms:
                    raise StencilLoweringError("boundary conditions are not affine")
                val = val - a * sym
                if a.free_symbols or float(a) != 0.0:
                    taps[key] = a if a.free_symbols else float(a)
            val = sp.expand(val)
            if val.free_symbols & fieldsyms:
                raise StencilLoweringError("boundary conditions are not affine")
            out[nd_] = (val, taps)
        return out

    # -- assemble ---------------------------------------------------------------------------------------------
    def lower(self) -> StencilProgram:
        eq_rpn = []
        for ev in range(self.nv):
            eq = self.eq_of[ev]
            resid = eq.lhs - eq.rhs
            dt_term = sp.Derivative(self.dvs[ev], self.t)
            # c Dt(u) + rest ~ 0 with a numeric c (Dt(u) ~ f gives c = 1, `v ~ Dt(u)` gives c = -1): the terms are lowered
            # in the residual AS WRITTEN -- the upwind direction is read off that form (array_discretization.jl:266-267)
            cdt = sp.expand(resid).coeff(dt_term)
            rest = sp.expand(resid) - cdt * dt_term if cdt != 1 else resid - dt_term
            if not cdt.is_number or cdt == 0 or rest.has(dt_term) or \
                    any(D.variables == (self.t,) for D in rest.atoms(sp.Derivative)):
                raise StencilLoweringError("equations must be of the form Dt(u) ~ f(...) (explicit ODE form)")
            self._point_reads = set()
            if self.segments is not None:
                for w_, dv in enumerate(self.dvs):
                    if rest.has(dv) and self.vax[w_][0] is not self.vax[ev][0] and self.vax[w_][0].n == 1:
                        # a variable of t alone (one chart node) read from the equation of a field: token s:w
                        if any(dv in D.atoms(sp.core.function.AppliedUndef) for D in rest.atoms(sp.Derivative)):
                            raise StencilLoweringError(f"{dv} inside a spatial derivative")
                        self._point_reads.add(w_)
                        continue
                    if rest.has(dv) and self.vax[w_][0] is not self.vax[ev][0]:
                        raise StencilLoweringError(f"{dv} appears in the equation of {self.dvs[ev]} but lives on another "
                                                   "domain: variables on different domains couple through interfaces only")
            ops = {}
            lowered = sum((self._lower_term(term, ops, ev) for term in self.split_additive(rest)), sp.Integer(0))
            eq_rpn.append(self.rpn(-lowered if cdt == 1 else -lowered / cdt, ops))
        self._point_reads = set()
        ghosts = self._ghosts()

        # core box: nodes where every node-indexed table of every equation is a core row
        # (a literal core on uniform axes; on non-uniform axes a "shape core": same taps, per-node weights from the table)
        corebox = None
        same_box = all(self.ilo[v] == self.ilo[0] and self.ihi[v] == self.ihi[0] for v in range(self.nv))
        if same_box:
            clo, chi = list(self.ilo[0]), list(self.ihi[0])
            ok = True
            node_tabs = {}
            for toks in eq_rpn:
                for tk in toks:
                    f = tk.split(":")
                    if f[0] == "L":
                        node_tabs.setdefault(int(f[3]), []).append(("L", int(f[1])))
                    elif f[0] == "N":
                        node_tabs.setdefault(int(f[2]), []).append(("N", int(f[4]), int(f[5]), int(f[6])))
                    elif f[0] == "W":
                        node_tabs.setdefault(int(f[3]), []).append(("W", int(f[1])))
                    elif f[0] == "M":                  # mixed derivatives run through the table-driven kernel only
                        ok = False
            for j, lst in node_tabs.items():
                for item in lst:
                    if item[0] == "L":
                        T = self.tabs[item[1]]
                        rng = T.core if T.core is not None else T.score
                        if rng is None:
                            ok = False; break
                        clo[j], chi[j] = max(clo[j], rng[0]), min(chi[j], rng[1])
                    elif item[0] == "W":
                        wid, lo, hi, rows = self.wtabs[item[1]]
                        # centre-target rows on five consecutive nodes: literal Jiang-Shu weights on a uniform axis,
                        # per-interval geometry arrays on a non-uniform one (built by the library at plan time)
                        core = _weno_core(rows)
                        if core is None:
                            ok = False; break
                        clo[j], chi[j] = max(clo[j], core[0]), min(chi[j], core[1])
                    else:
                        TI, TD, TO = self.tabs[item[1]], self.tabs[item[2]], self.tabs[item[3]]
                        if TI.core is None or TD.core is None or TO.core is None:
                            ok = False; break
                        olo, ohi, ooff, ow = TO.core
                        # half points tapped by node i: i+ooff .. i+ooff+L-1 must be core rows of TI and TD
                        mlo = max(TI.core[0], TD.core[0]); mhi = min(TI.core[1], TD.core[1])
                        clo[j] = max(clo[j], olo, mlo - ooff)
                        chi[j] = min(chi[j], ohi, mhi - ooff - (TO.L - 1))
                if not ok:
                    break
            if ok and all(chi[j] >= clo[j] for j in range(self.nd)):
                corebox = (clo, chi)

        out = ["MOLPROG 1", f"ndim {self.nd}", f"nvar {self.nv}", f"nparam {len(self.params)}"]
        for k, (p, val) in enumerate(zip(self.params, self.pvals)):
            out.append(f"param {k} {p} {_hex(val)}")
        for j, ax in enumerate(self.axes):
            out.append(f"grid {j} {ax.n} {'U' if ax.uniform else 'N'} {_hex(ax.dx if ax.uniform else 0.0)}")
            out.append(f"coords {j} " + " ".join(_hex(v) for v in ax.x))
        for v in range(self.nv):
            out.append(f"var {v} {self.fns[v]}")
            out.append(f"interior {v} " + " ".join(map(str, self.ilo[v] + self.ihi[v])))
            out.append(f"periodic {v} " + " ".join(str(int(b)) for b in self.per[v]))
        for T in self.tabs:
            T.emit(out)
        for wid, lo, hi, rows in self.wtabs:
            out.append(f"wtab {wid} {hi - lo + 1} {lo}")
            core = _weno_core(rows)
            if core:
                out.append(f"wcore {wid} {core[0]} {core[1]}")
            for i in sorted(rows):
                if core and core[0] <= i <= core[1]:
                    continue
                out.append(f"wrow {wid} {i} {rows[i][0]} {rows[i][1]}")
        for fid, toks in enumerate(self.fn_exprs):
            out.append(f"fn {fid} {len(toks)} " + " ".join(toks))
        for (v, j, node), (G, taps) in sorted(ghosts.items()):
            toks = self.rpn(G, allow_fields=False)
            if all(isinstance(a, float) for a in taps.values()):
                tl = " ".join(f"{w_} {tp} {_hex(a)}" for (w_, tp), a in sorted(taps.items()))
                out.append(f"ghost {v} {j} {node} {len(taps)} {tl} {len(toks)} " + " ".join(toks))
            else:       # expression coefficients: ghostx v dim node ntaps nG G.. (var node nc coef..)*
                parts = [f"ghostx {v} {j} {node} {len(taps)} {len(toks)}"] + toks
                for (w_, tp), a in sorted(taps.items(), key=lambda kv: kv[0]):
                    ct = self.rpn(sp.sympify(a), allow_fields=False)
                    parts += [str(w_), str(tp), str(len(ct))] + ct
                out.append(" ".join(parts))
        for v, toks in enumerate(eq_rpn):
            out.append(f"eq {v} {len(toks)} " + " ".join(toks))
        if corebox is not None:
            out.append("corebox " + " ".join(map(str, corebox[0] + corebox[1])))
        out.append("end")

        P = StencilProgram()
        P.text = "\n".join(out) + "\n"
        P.var_names = [str(f) for f in self.fns]
        P.ilo, P.ihi = self.ilo, self.ihi
        P.shapes = [tuple(self.ihi[v][j] - self.ilo[v][j] + 1 for j in range(self.nd)) for v in range(self.nv)]
        sizes = [int(np.prod(s)) for s in P.shapes]
        P.offsets = [int(o) for o in np.concatenate([[0], np.cumsum(sizes)])[:-1]]
        P.nstate = int(sum(sizes))
        P.axes = self.axes
        P.tspan = self.tspan
        P.params, P.pvals = self.params, self.pvals
        P.periodic = self.per
        P.segments = self.segments
        P.corebox = corebox
        P._u0_fn = lambda: self._initial(P)
        return P

    # -- initial condition at the interior nodes (generate_ic_defaults.jl:11-21) ---------------------------------
    def _initial(self, P):
        u0 = np.zeros(P.nstate)
        for v in range(self.nv):
            if self.ic[v] is None:
                raise StencilLoweringError(f"missing initial condition for {self.dvs[v]}")
            coords = []
            for j in range(self.nd):
                shape = [1] * self.nd
                g = self.axes[j].x[self.ilo[v][j] - 1:self.ihi[v][j]]
                shape[j] = len(g)
                coords.append(g.reshape(shape))
            f = sp.lambdify(self.xs + [self.t] + self.params, self.ic[v], "numpy")
            val = f(*coords, self.tspan[0], *self.pvals)
            val = np.broadcast_to(np.asarray(val, dtype=float), P.shapes[v])
            u0[P.offsets[v]:P.offsets[v] + val.size] = val.ravel(order="F")
        return u0


def lower(pdesys, disc) -> StencilProgram:
    return Lowering(pdesys, disc).lower()
