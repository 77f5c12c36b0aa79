"""Oracle for variables on DIFFERENT 1-D domains joined by interface boundary conditions.

Test infrastructure only (see oracle/__init__).  Restates, node by node, what the reference's scalarized
discretization does for systems such as

    Dt(u1(t, x1)) ~ -v Dx1(u1(t, x1)),   Dt(u2(t, x2)) ~ -v Dx2(u2(t, x2)),   u1(t, b) ~ u2(t, b)

(test/Diffusion/MOL_1D_Linear_Diffusion.jl:887-930, test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl:122-215,
test/Convection_WENO/MOL_1D_WENO_NU_Interface.jl).  Every variable keeps its OWN grid and its own full-grid array
(discretize_vars.jl:256-273); a stencil tap that leaves the grid through an interface end is redirected onto the
neighbouring variable's array exactly as `_wrapinterface` does (interface_boundary.jl:79-107: index + (l2 - 1) through a
lower interface, index + 1 - l1 through an upper one), and its coordinate is read in the differentiated variable's chart
as `bcoord` does (:109-153).  This is deliberately a different construction from the lowering's single chart axis.

Followed here:
  interior / extents    interior_map.jl:1-10 (an interface clips its LOWER end only), :117-139 (no extent at an open end)
  interface alias       generate_bc_eqs.jl:35-58   u2[1] ~ u1[l1]
  centered              centered_difference.jl:5-57 (uniform grids only across interfaces, interior_map.jl:33-52)
  upwind                upwind_difference.jl:1-28 (uniform), :85-199 (non-uniform: weights computed on the spot from chart
                        coordinates where the stencil crosses an interface or sits on the table's blind spot)
  WENO                  function_scheme.jl:1-76 with bcoord coordinates, WENO.jl / nonuniform_weno.jl kernels
  boundary solves       generate_bc_eqs.jl:238-328 (affine in the edge node)
"""
import numpy as np
import sympy as sp

from . import operators as ops
from . import weno as wk
from .discretize import make_grid
from .evalexpr import evaluate, evaluate_abs
from .fornberg import calculate_weights


class InterfaceOracle1D:
    def __init__(self, pdesys, disc):
        self.sys, self.disc = pdesys, disc
        self.t = disc.time
        self.dvs = list(pdesys.dvs)
        self.funcs = [d.func for d in self.dvs]
        self.nv = len(self.dvs)
        self.xv = []
        for d in self.dvs:
            sa = [a for a in d.args if a != self.t]
            assert len(sa) <= 1, "oracle scope: interface systems in one spatial dimension"
            # a variable of t alone (an ODE next to the PDEs): a single node, no spatial operators
            self.xv.append(sa[0] if sa else sp.Symbol(f"__point_{d.func}"))
        self.tpos = [[k for k, a in enumerate(d.args) if a == self.t][0] for d in self.dvs]
        self.dom = {iv.var: (float(iv.lo), float(iv.hi)) for iv in pdesys.domains}
        self.tspan = self.dom[self.t]
        self.params = [p for p, _ in pdesys.ps]
        self.pvals = np.array([v for _, v in pdesys.ps], dtype=float)
        assert type(disc.grid_align).__name__ == "CenterAlignedGrid"
        self.grid, self.dx = [], []
        for x in self.xv:
            if x not in self.dom:
                self.grid.append(np.array([0.0])); self.dx.append(None)
                continue
            g, dx = make_grid(self.dom[x][0], self.dom[x][1], disc.dxs[x])
            self.grid.append(g)
            self.dx.append(dx)
        self.n = [len(g) for g in self.grid]
        self.weno = type(disc.advection_scheme).__name__ == "WENOScheme"
        self.weno_eps = getattr(disc.advection_scheme, "epsilon", None)
        self.upwind_order = getattr(disc.advection_scheme, "order", 1)
        self._parse_bcs()
        self._operators()
        self._interiors()
        self._initial()

    # -- boundary conditions ------------------------------------------------------------------------------------------
    def _at(self, val, x, upper):
        ref = self.dom[x][1 if upper else 0]
        return abs(float(val) - ref) <= 1e-12 * max(1.0, abs(ref))

    def _parse_bcs(self):
        self.lower_nb = [None] * self.nv      # variable joined at the lower end (InterfaceBoundary{lower})
        self.upper_nb = [None] * self.nv
        self.bounds = [[None, None] for _ in range(self.nv)]
        self.ics = [None] * self.nv
        for bc in self.sys.bcs:
            L, R = bc.lhs, bc.rhs
            fl, fr = getattr(L, "func", None), getattr(R, "func", None)
            if fl in self.funcs and fr in self.funcs and fl != fr:
                a, b = self.funcs.index(fl), self.funcs.index(fr)
                va = L.args[1 - self.tpos[a]] if len(L.args) > 1 else sp.Symbol("__none")
                vb = R.args[1 - self.tpos[b]] if len(R.args) > 1 else sp.Symbol("__none")
                if va.is_number and vb.is_number:
                    if self._at(va, self.xv[a], True) and self._at(vb, self.xv[b], False):
                        lo_v, up_v = a, b
                    elif self._at(va, self.xv[a], False) and self._at(vb, self.xv[b], True):
                        lo_v, up_v = b, a
                    else:
                        raise ValueError(f"interface {bc} joins two variables at the same end")
                    self.upper_nb[lo_v], self.lower_nb[up_v] = up_v, lo_v
                    continue
            found = False
            for v, fn in enumerate(self.funcs):
                calls = [c for c in (L - R).atoms(sp.core.function.AppliedUndef) if c.func == fn]
                for call in calls:
                    num = [(k, a) for k, a in enumerate(call.args) if a.is_number]
                    if not num:
                        continue
                    k, val = num[0]
                    if k == self.tpos[v]:
                        assert L == call, "initial condition must read u(t0, x) ~ expr"
                        self.ics[v] = R
                    else:
                        upper = self._at(val, self.xv[v], True)
                        assert upper or self._at(val, self.xv[v], False), bc
                        self.bounds[v][int(upper)] = bc
                    found = True
                    break
                if found:
                    break
            assert found, f"could not classify {bc}"

    def _operators(self):
        exprs = [e.lhs - e.rhs for e in self.sys.eqs] + [b.lhs - b.rhs for b in self.sys.bcs]
        self.dd = []
        for v in range(self.nv):
            if len(self.grid[v]) == 1:
                self.dd.append(None)
                continue
            orders = set()
            for e in exprs:
                for D in e.atoms(sp.Derivative):
                    for var, cnt in D.variable_count:
                        if var == self.xv[v]:
                            orders.add(int(cnt))
            self.dd.append(ops.differential_discretizer(self.grid[v], self.dx[v], sorted(orders), self.disc.approx_order,
                                                        self.upwind_order, self.weno))
        # validate_interface_orders (interior_map.jl:33-52)
        for v in range(self.nv):
            w_ = self.upper_nb[v]
            if w_ is None:
                continue
            same = self.dx[v] is not None and self.dx[w_] is not None and abs(self.dx[v] - self.dx[w_]) <= 1e-12 * self.dx[v]
            if not same:
                for vv in (v, w_):
                    assert all(d <= 1 for d in self.dd[vv].orders), \
                        "only first-order derivatives are supported across mismatched grids"

    def _eqvar(self, eq):
        for D in (eq.lhs - eq.rhs).atoms(sp.Derivative):
            if D.variables == (self.t,) and D.expr in self.dvs:
                return self.dvs.index(D.expr)
        raise ValueError(eq)

    def _interiors(self):
        self.eq_of_var = {self._eqvar(eq): eq for eq in self.sys.eqs}
        assert len(self.eq_of_var) == self.nv
        self.ilo, self.ihi, self.vlower, self.vupper, self.ext = [], [], [], [], []
        for v in range(self.nv):
            haslower, hasupper = self.lower_nb[v] is not None, self.upper_nb[v] is not None
            lo = int(self.bounds[v][0] is not None) + int(haslower)          # clip_interior!!: interface clips its lower end
            up = int(self.bounds[v][1] is not None)
            resid = self.eq_of_var[v].lhs - self.eq_of_var[v].rhs
            orders = {int(c) for D in resid.atoms(sp.Derivative) for var, c in D.variable_count if var == self.xv[v]}
            le = ue = 0
            for d in orders:
                if d % 2 == 1:
                    e = 2 if (d == 1 and self.weno and self.dx[v] is not None) else 0
                    if not haslower:
                        le = max(le, e)
                    if not hasupper:
                        ue = max(ue, e)
            self.vlower.append(lo); self.vupper.append(up); self.ext.append((le, ue))
            self.ilo.append(1 + max(lo, le))
            self.ihi.append(self.n[v] - max(up, ue))
        self._base = np.concatenate([[0], np.cumsum(self.n)]).astype(int)       # offsets of the full-grid arrays, concatenated
        self._cache = {}
        self.sizes = [self.ihi[v] - self.ilo[v] + 1 for v in range(self.nv)]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(int)
        self.nstate = int(self.offsets[-1])

    def _env(self, v, x, t, p, own=None):
        env = {self.xv[v]: x, self.t: t}
        env.update({s: float(q) for s, q in zip(self.params, p)})
        if own is not None:
            env[self.dvs[v]] = own
        return env

    def _initial(self):
        u0 = np.zeros(self.nstate)
        for v in range(self.nv):
            x = self.grid[v][self.ilo[v] - 1:self.ihi[v]]
            val = evaluate(self.ics[v], self._env(v, x, self.tspan[0], self.pvals))
            u0[self.offsets[v]:self.offsets[v + 1]] = np.broadcast_to(np.asarray(val, dtype=float), x.shape)
        self.u0 = u0

    # -- taps across interfaces (interface_boundary.jl:79-153) -------------------------------------------------------
    def wrap(self, v, i):
        """bwrap: (variable, index) a raw tap index of variable v refers to."""
        if i <= 1 and self.lower_nb[v] is not None:
            w_ = self.lower_nb[v]
            return w_, i + (self.n[w_] - 1)
        if i > self.n[v] and self.upper_nb[v] is not None:
            return self.upper_nb[v], i + 1 - self.n[v]
        return v, i

    def bcoord(self, v, i):
        """Coordinate of raw tap index i in v's chart."""
        if i <= 1 and self.lower_nb[v] is not None:
            g1, g2 = self.grid[v], self.grid[self.lower_nb[v]]
            return g2[i + len(g2) - 1 - 1] - (g2[-1] - g1[0])
        if i > self.n[v] and self.upper_nb[v] is not None:
            g1, g2 = self.grid[v], self.grid[self.upper_nb[v]]
            return g2[i + 1 - len(g1) - 1] + (g1[-1] - g2[0])
        return self.grid[v][i - 1]

    # -- per-node stencil rows: (weights, raw tap indices of variable v) ---------------------------------------------------
    @staticmethod
    def _row(w, raw):
        return np.asarray(w, dtype=float), list(raw)

    def centered(self, full, v, d, i):
        D = self.dd[v].map[d]
        n = self.n[v]
        haslower, hasupper = self.lower_nb[v] is not None, self.upper_nb[v] is not None
        bpc, bsl, L = D.boundary_point_count, D.boundary_stencil_length, D.stencil_length
        if D.uniform:
            if i <= bpc and not haslower:
                return self._row(D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)])
            if i > n - bpc and not hasupper:
                return self._row(D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)])
            return self._row(D.stencil_coefs, [i + k for k in range(-(L // 2), L // 2 + 1)])
        assert not (haslower or hasupper), "centered differences across interfaces need uniform grids"
        if i <= bpc:
            return self._row(D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)])
        if i > n - bpc:
            return self._row(D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)])
        return self._row(D.stencil_coefs[i - bpc - 1], [i + k for k in range(-(L // 2), L // 2 + 1)])

    def upwind(self, full, v, d, i, ispositive):
        D = (self.dd[v].windneg if ispositive else self.dd[v].windpos)[d]
        n = self.n[v]
        haslower, hasupper = self.lower_nb[v] is not None, self.upper_nb[v] is not None
        L, bsl = D.stencil_length, D.boundary_stencil_length
        if D.uniform:                                                   # upwind_difference.jl:1-28
            if not ispositive:
                if i > n - D.boundary_point_count and not hasupper:
                    return self._row(D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)])
                return self._row(D.stencil_coefs, [i + k for k in range(L)])
            if i <= D.offside and not haslower:
                return self._row(D.low_boundary_coefs[i - 1], [1 + k for k in range(bsl)])
            return self._row(D.stencil_coefs, [i + k for k in range(-L + 1, 1)])
        offsets = range(0, L) if not ispositive else range(-L + 1, 1)
        if haslower or hasupper:                                        # :85-129
            raw = [i + k for k in offsets]
            crossing = any(r < 1 or r > n for r in raw)
            blind = (i == n) if ispositive else (i == 1)
            if crossing or blind:
                for r in raw:
                    w_, k = self.wrap(v, r)
                    assert 1 <= k <= self.n[w_], "upwind stencil extends past a non-interface boundary"
                w = calculate_weights(d, self.bcoord(v, i), [self.bcoord(v, r) for r in raw])
                return self._row(w, raw)
        if not ispositive:                                              # :163-199
            if i > n - D.boundary_point_count:
                return self._row(D.high_boundary_coefs[n - i], [n - bsl + 1 + k for k in range(bsl)])
            return self._row(D.stencil_coefs[i - 1], [i + k for k in range(L)])
        return self._row(D.stencil_coefs[i - D.offside - 1], [i + k for k in range(-L + 1, 1)])

    def weno_taps(self, v, i):
        """function_scheme.jl:1-34: reconstruction target T, raw taps and their chart coordinates (None: uniform)."""
        n = self.n[v]
        haslower, hasupper = self.lower_nb[v] is not None, self.upper_nb[v] is not None
        if i <= 2 and not haslower:
            T, raw = i, [1 + k for k in range(5)]
        elif i > n - 2 and not hasupper:
            T, raw = 5 - (n - i), [n - 4 + k for k in range(5)]
        else:
            assert not ((haslower or hasupper) and n - 1 < 5)
            T, raw = 3, [i + k for k in range(-2, 3)]
        if self.dx[v] is not None:
            assert T == 3, "uniform WENO is only defined on the interior"
            return T, raw, None
        xx = [self.bcoord(v, r) for r in raw] if (haslower or hasupper) else [self.grid[v][r - 1] for r in raw]
        return T, raw, xx

    # -- assembly over the interior of one equation (rows do not depend on u: built once, cached) ----------------------------
    def _gidx(self, v, r):
        w_, k = self.wrap(v, r)
        assert 1 <= k <= self.n[w_], (v, r, w_, k)
        return self._base[w_] + k - 1

    def _linear(self, key, rowfn, ev, full):
        import scipy.sparse as sps
        if key not in self._cache:
            ri, ci, vv = [], [], []
            for r_, i in enumerate(range(self.ilo[ev], self.ihi[ev] + 1)):
                w, raw = rowfn(i)
                for wk_, r in zip(w, raw):
                    ri.append(r_); ci.append(self._gidx(ev, r)); vv.append(float(wk_))
            self._cache[key] = sps.csr_matrix((vv, (ri, ci)), shape=(self.sizes[ev], int(self._base[-1])))
        M = self._cache[key]
        cat = np.concatenate(full)
        return (abs(M) @ np.abs(cat)) if self._absmode else (M @ cat)

    def _weno(self, ev, full):
        key = ("weno", ev)
        if key not in self._cache:
            groups = {}
            for r_, i in enumerate(range(self.ilo[ev], self.ihi[ev] + 1)):
                T, raw, xx = self.weno_taps(ev, i)
                g = groups.setdefault(T, ([], [], []))
                g[0].append(r_); g[1].append([self._gidx(ev, r) for r in raw]); g[2].append(xx)
            self._cache[key] = {T: (np.array(a), np.array(b), None if c[0] is None else np.array(c, dtype=float))
                                for T, (a, b, c) in groups.items()}
        cat = np.concatenate(full)
        out = np.zeros(self.sizes[ev])
        for T, (rows, taps, xx) in self._cache[key].items():
            uu = [cat[taps[:, k]] for k in range(5)]
            if xx is None:
                out[rows] = wk.weno_f_uniform(uu, self.weno_eps, self.dx[ev])
            else:
                out[rows] = wk.weno_f_nonuniform_core(uu, self.weno_eps, [xx[:, k] for k in range(5)], T)
        return np.abs(out) if self._absmode else out

    # -- boundary fill --------------------------------------------------------------------------------------------------
    def unpack(self, u):
        full = []
        for v in range(self.nv):
            U = np.zeros(self.n[v])
            U[self.ilo[v] - 1:self.ihi[v]] = u[self.offsets[v]:self.offsets[v + 1]]
            full.append(U)
        return full

    def _fill_boundaries(self, full, t, p):
        # interface alias: the variable whose LOWER end is an interface copies its neighbour's upper edge node
        for v in range(self.nv):
            if self.lower_nb[v] is not None:
                w_ = self.lower_nb[v]
                full[v][0] = full[w_][self.n[w_] - 1]
        for v in range(self.nv):
            if self.dd[v] is None:
                continue
            for side in (0, 1):
                bc = self.bounds[v][side]
                pads = self._pads(v, bool(side))
                if bc is None:
                    assert not pads
                    continue
                # The boundary equation and the extrapolation equations of the pad nodes next to it (generate_extrap_eqs!,
                # generate_bc_eqs.jl:336-392) are algebraic equations of one linear system (a Neumann row reads the pad,
                # the pad's row reads the edge node): solved together by probing the affine residuals.
                nodes = [self.n[v] if side else 1] + pads
                U = full[v]

                def resid(vals, v=v, side=side, bc=bc, pads=pads, nodes=nodes, U=U):
                    W = U.copy()
                    for nd_, val in zip(nodes, vals):
                        W[nd_ - 1] = val
                    out = [self._bc_residual(W, v, side, bc, t, p)]
                    for pad in pads:
                        w, taps = self._pad_row(v, pad)
                        out.append(W[pad - 1] - sum(wk_ * W[tp - 1] for wk_, tp in zip(w, taps)))
                    return np.array(out, dtype=float)
                m = len(nodes)
                F0 = resid([0.0] * m)
                A = np.stack([resid([1.0 if i == k else 0.0 for i in range(m)]) - F0 for k in range(m)], axis=1)
                sol = np.linalg.solve(A, -F0)
                for nd_, val in zip(nodes, sol):
                    U[nd_ - 1] = val

    def _pads(self, v, upper):
        le, ue = self.ext[v]
        n = self.n[v]
        e, vl = (ue, self.vupper[v]) if upper else (le, self.vlower[v])
        if upper and self.upper_nb[v] is not None or (not upper and self.lower_nb[v] is not None):
            return []
        out, ninterp = [], e - vl
        while ninterp >= vl and vl > 0:
            node = (n - ninterp) if upper else (1 + ninterp)
            ninterp -= 1
            if not (self.ilo[v] <= node <= self.ihi[v]):
                out.append(node)
        return out

    def _pad_row(self, v, node):
        B, n = self.dd[v].boundary, self.n[v]
        bsl = B.boundary_stencil_length
        if node <= B.boundary_point_count:
            return B.low_boundary_coefs[node - 1], [1 + k for k in range(bsl)]
        return B.high_boundary_coefs[n - node], [n - bsl + 1 + k for k in range(bsl)]

    def _bc_residual(self, W, v, side, bc, t, p):
        """lhs - rhs of a boundary condition on the full-grid array W (generate_bc_eqs.jl:238-328: u(t, x_b) -> the edge
        node, Dx^d u(t, x_b) -> the one-sided row of the centred operator at the edge node).  The symbolic form (taps as
        placeholder symbols) is built once per boundary."""
        cache = self.__dict__.setdefault("_bc_cache", {})
        if (v, side) not in cache:
            n = self.n[v]
            node = n if side else 1
            resid = bc.lhs - bc.rhs
            taps_of = {}

            def tap(tp):
                return taps_of.setdefault(tp, sp.Symbol(f"__w{tp}"))
            subs = {}
            for D in resid.atoms(sp.Derivative):
                assert D.expr.func == self.funcs[v], f"unsupported BC derivative {D}"
                (var, cnt), = D.variable_count
                Dop = self.dd[v].map[int(cnt)]
                bsl = Dop.boundary_stencil_length
                if side:
                    w, taps = Dop.high_boundary_coefs[0], [n - bsl + 1 + k for k in range(bsl)]
                else:
                    w, taps = Dop.low_boundary_coefs[0], [1 + k for k in range(bsl)]
                subs[D] = sum(float(wk_) * tap(tp) for wk_, tp in zip(w, taps))
            resid = resid.xreplace(subs)
            for call in [c for c in resid.atoms(sp.core.function.AppliedUndef) if c.func == self.funcs[v]]:
                resid = resid.xreplace({call: tap(node)})
            cache[(v, side)] = (resid, taps_of, node)
        resid, taps_of, node = cache[(v, side)]
        env = self._env(v, self.grid[v][node - 1], t, p)
        env.update({s_: float(W[tp - 1]) for tp, s_ in taps_of.items()})
        return float(evaluate(resid, env))

    def full_state(self, u, t, p=None):
        p = self.pvals if p is None else np.asarray(p, dtype=float)
        full = self.unpack(np.asarray(u, dtype=float))
        self._fill_boundaries(full, t, p)
        return full

    # -- the RHS ---------------------------------------------------------------------------------------------------------
    _absmode = False

    @staticmethod
    def split_additive(expr):
        out = []
        for term in sp.Add.make_args(expr):
            c, rest = term.as_coeff_Mul()
            if isinstance(rest, sp.Add):
                out += [c * q for q in InterfaceOracle1D.split_additive(rest)]
            else:
                out.append(term)
        return out

    def _lower_term(self, term, ev, ph):
        """Derivatives -> placeholder symbols bound to thunks full -> array (the structure is state-independent)."""
        full = None                                   # rows are built from grids only; `full` is unused by the row functions

        def new(thunk):
            s = sp.Symbol(f"__d{len(ph)}")
            ph[s] = thunk
            return s
        factors = list(sp.Mul.make_args(term))
        for k, fct in enumerate(factors):
            if isinstance(fct, sp.Derivative) and fct.expr in self.dvs and len(fct.variable_count) == 1:
                x, d = fct.variable_count[0]
                d = int(d)
                u = self.dvs.index(fct.expr)
                if d % 2 == 1 and not (self.weno and d == 1) and len(factors) > 1:
                    assert u == ev and x == self.xv[u]
                    coef = sp.Mul(*(factors[:k] + factors[k + 1:]))
                    assert not coef.atoms(sp.Derivative)
                    bwd = new(lambda F, u=u, d=d: self._linear(("u", u, d, True), lambda i: self.upwind(None, u, d, i, True), ev, F))
                    fwd = new(lambda F, u=u, d=d: self._linear(("u", u, d, False), lambda i: self.upwind(None, u, d, i, False), ev, F))
                    return sp.Piecewise((coef * bwd, coef > 0), (coef * fwd, True))
        subs = {}
        for D in term.atoms(sp.Derivative):
            assert D.expr in self.dvs and len(D.variable_count) == 1, f"unsupported derivative {D}"
            x, d = D.variable_count[0]
            d = int(d)
            u = self.dvs.index(D.expr)
            assert u == ev and x == self.xv[u], "oracle scope: derivatives of the equation's own variable"
            if d % 2 == 0:
                subs[D] = new(lambda F, u=u, d=d: self._linear(("c", u, d), lambda i: self.centered(None, u, d, i), ev, F))
            elif self.weno and d == 1:
                subs[D] = new(lambda F: self._weno(ev, F))
            else:
                subs[D] = new(lambda F, u=u, d=d: self._linear(("u", u, d, True), lambda i: self.upwind(None, u, d, i, True), ev, F))
        return term.xreplace(subs)

    def rhs_termscale(self, u, t, p=None):
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
            eq = self.eq_of_var[ev]
            resid = eq.lhs - eq.rhs
            rest = resid - sp.Derivative(self.dvs[ev], self.t)
            if ("eq", ev) not in self._cache:
                ph = {}
                lowered = sum(self._lower_term(term, ev, ph) for term in self.split_additive(rest))
                self._cache[("eq", ev)] = (sp.sympify(lowered), ph)
            lowered, thunks = self._cache[("eq", ev)]
            x = self.grid[ev][self.ilo[ev] - 1:self.ihi[ev]]
            env = self._env(ev, x, t, p, full[ev][self.ilo[ev] - 1:self.ihi[ev]])
            for w_ in range(self.nv):           # variables of t alone are visible to every equation (their single node)
                if w_ != ev and len(self.grid[w_]) == 1 and self.xv[w_] not in self.dom:
                    env[self.dvs[w_]] = float(full[w_][0])
            env.update({s_: th(full) for s_, th in thunks.items()})
            val = (evaluate_abs if self._absmode else evaluate)(lowered, env)
            val = (1.0 if self._absmode else -1.0) * np.broadcast_to(np.asarray(val, dtype=float), x.shape)
            du[self.offsets[ev]:self.offsets[ev + 1]] = val
        return du
