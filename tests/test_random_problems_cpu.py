"""Seeded random 1-D problems: random grid (uniform / random node vector), approx_order 2 / 4 / 6, UpwindScheme or WENOScheme,
a random mix of diffusion, linear or nonlinear advection and a reaction / source term, and an independent random choice of
Dirichlet / Neumann / Robin data at each end.  The lowering's stencil program (executed by tests/ir_interp.py) must
reproduce the oracle's du on every one of them -- this is how the uniform-WENO + Neumann case (a derivative condition
next to an extrapolation pad: two coupled algebraic equations) was found and is kept fixed."""
import numpy as np
import pytest
import sympy as sp

import mol_b200
from mol_b200.interface import Differential, Eq, Interval, MOLFiniteDifference, PDESystem, UpwindScheme, WENOScheme
from oracle.discretize import OracleProblem

from ir_interp import IRProgram


def random_problem(rng):
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    U = u(t, x)
    Dt, Dx = Differential(t), Differential(x)
    order = int(rng.choice([2, 4, 6]))
    n = int(rng.integers(14, 30))
    if rng.integers(2):
        g = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0.03, 0.97, n - 2)]))
        if np.diff(g).min() < 1e-3:
            g = np.linspace(0.0, 1.0, n) ** 1.2
    else:
        g = 1.0 / (n - 1)
    weno = bool(rng.integers(2))
    terms = []
    if rng.integers(2):
        terms.append(float(rng.uniform(0.1, 2)) * (Dx ** 2)(U))
    adv = int(rng.integers(3))
    if adv == 1:
        terms.append(-float(rng.uniform(-1, 1)) * Dx(U))
    elif adv == 2:
        terms.append(-U * Dx(U))
    if rng.integers(2):
        terms.append(U * (1 - U) + sp.sin(x) * sp.exp(-t))
    if not terms:
        terms.append((Dx ** 2)(U))
    bcs = [Eq(u(0, x), sp.cos(2 * x) + 1.5)]
    kinds = []
    for end in (0.0, 1.0):
        kind = str(rng.choice(["dirichlet", "neumann", "robin"]))
        kinds.append(kind)
        if kind == "dirichlet":
            bcs.append(Eq(u(t, end), sp.exp(-t) * 1.3))
        elif kind == "neumann":
            bcs.append(Eq(Dx(u(t, end)), 0.2 * sp.exp(-t)))
        else:
            bcs.append(Eq(Dx(u(t, end)) + float(rng.uniform(0.5, 2)) * u(t, end), sp.cos(t)))
    sys_ = PDESystem([Eq(Dt(U), sum(terms))], bcs, [Interval(t, 0.0, 1.0), Interval(x, 0.0, 1.0)], [t, x], [U])
    disc = MOLFiniteDifference({x: g}, t, approx_order=order, advection_scheme=WENOScheme() if weno else UpwindScheme())
    return sys_, disc, f"order {order}, {'vector' if np.ndim(g) else 'uniform'} grid n = {n}, {'WENO' if weno else 'upwind'}, {kinds}, {terms}"


@pytest.mark.parametrize("seed", range(6))
def test_stencil_program_matches_oracle_on_random_problems(seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(8):
        sys_, disc, what = random_problem(rng)
        prog = mol_b200.symbolic_discretize(sys_, disc)
        orc = OracleProblem(sys_, disc)
        ir = IRProgram(prog.text)
        assert prog.nstate == orc.nstate, what
        u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        for tt in (0.0, 0.37):
            ref, got = orc.rhs(u, tt), ir.rhs(u, tt)
            scale = float(np.max(orc.rhs_termscale(u, tt)))
            assert float(np.max(np.abs(ref - got))) <= 1e-12 * scale, what


def random_system_2d(rng, nmin=14, nmax=26):
    """One or two variables on [0, 1]^2: Laplacian, linear or nonlinear (cross-variable) advection, reaction + source, each
    optional; per dimension a uniform or power-law node vector, periodic with probability 1/4, else independent Dirichlet /
    Neumann / Robin data (varying along the wall) per wall and variable; approx_order 2 / 4; WENO with probability 1/3."""
    t, x, y = sp.symbols("t x y")
    fs = [sp.Function("u"), sp.Function("v")][:int(rng.integers(1, 3))]
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    order = int(rng.choice([2, 4]))
    grids = {}
    for s_ in (x, y):
        n = int(rng.integers(nmin, nmax))
        if rng.integers(3) == 0:
            g = np.linspace(0, 1, n) ** float(rng.uniform(1.0, 1.4))
            g[-1] = 1.0
            grids[s_] = g
        else:
            grids[s_] = 1.0 / (n - 1)
    weno = bool(rng.integers(3) == 0)
    per = {x: bool(rng.integers(4) == 0), y: bool(rng.integers(4) == 0)}
    eqs, bcs = [], []
    for k, f in enumerate(fs):
        F = f(t, x, y)
        terms = [float(rng.uniform(0.1, 1)) * ((Dx ** 2)(F) + (Dy ** 2)(F))] if rng.integers(4) else []
        a = int(rng.integers(3))
        if a == 1:
            terms.append(-float(rng.uniform(-1, 1)) * Dx(F) - float(rng.uniform(-1, 1)) * Dy(F))
        if a == 2:
            terms.append(-fs[0](t, x, y) * Dx(F) - fs[-1](t, x, y) * Dy(F))
        if rng.integers(2):
            terms.append(F * (1 - fs[-1](t, x, y)) + sp.sin(x + y) * sp.exp(-t))
        if not terms:
            terms.append((Dx ** 2)(F))
        eqs.append(Eq(Dt(F), sum(terms)))
        bcs.append(Eq(f(0, x, y), sp.cos(2 * x + k) * sp.sin(y + 0.3) + 1.5))
        for s_, D_ in ((x, Dx), (y, Dy)):
            at = (lambda val, f=f: f(t, val, y)) if s_ == x else (lambda val, f=f: f(t, x, val))
            other = y if s_ == x else x
            if per[s_]:
                bcs.append(Eq(at(0.0), at(1.0)))
                continue
            for end in (0.0, 1.0):
                kind = str(rng.choice(["dirichlet", "neumann", "robin"]))
                if kind == "dirichlet":
                    bcs.append(Eq(at(end), sp.exp(-t) * (1.3 + other)))
                elif kind == "neumann":
                    bcs.append(Eq(D_(at(end)), 0.2 * sp.exp(-t) * other))
                else:
                    bcs.append(Eq(D_(at(end)) + float(rng.uniform(0.5, 2)) * at(end), sp.cos(t) + other))
    sys_ = PDESystem(eqs, bcs, [Interval(t, 0.0, 1.0), Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0)], [t, x, y],
                     [f(t, x, y) for f in fs])
    disc = MOLFiniteDifference(grids, t, approx_order=order, advection_scheme=WENOScheme() if weno else UpwindScheme())
    return sys_, disc, f"order {order}, WENO {weno}, periodic {[per[x], per[y]]}, {[str(e) for e in eqs]}, {[str(b) for b in bcs]}"


@pytest.mark.parametrize("seed", range(3))
def test_stencil_program_matches_oracle_on_random_2d_systems(seed):
    from mol_b200.lowering import StencilLoweringError
    rng = np.random.default_rng(2000 + seed)
    done = 0
    while done < 6:
        sys_, disc, what = random_system_2d(rng)
        try:
            prog = mol_b200.symbolic_discretize(sys_, disc)
        except StencilLoweringError as e:           # what the reference rejects too (periodic + non-uniform centred rows)
            assert "non-uniform grids for centered" in str(e), what
            continue
        done += 1
        orc = OracleProblem(sys_, disc)
        ir = IRProgram(prog.text)
        u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        for tt in (0.0, 0.37):
            ref, got = orc.rhs(u, tt), ir.rhs(u, tt)
            assert float(np.max(np.abs(ref - got))) <= 1e-12 * float(np.max(orc.rhs_termscale(u, tt))), what


def test_generated_kernels_match_oracle_on_random_2d_systems():
    """The same kind of random systems, large enough for several tiles, through the emulated generated kernels
    (tests/cuda_emu): table-driven on every unknown, tiled (cooperative loader) on its core box."""
    from cuda_emu import EmuKernel
    from mol_b200 import capi
    from mol_b200.lowering import StencilLoweringError
    from test_generated_code_cpu import _core_mask
    rng = np.random.default_rng(3000)
    done = tiled = 0
    while done < 3:
        sys_, disc, what = random_system_2d(rng, 40, 80)
        try:
            prog = mol_b200.symbolic_discretize(sys_, disc)
        except StencilLoweringError:
            continue
        done += 1
        orc = OracleProblem(sys_, disc)
        plan = capi.Plan(prog.text, device=-1)
        u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        ref = orc.rhs(u, 0.37)
        scale = float(np.max(orc.rhs_termscale(u, 0.37)))
        assert float(np.max(np.abs(EmuKernel(plan, prog).rhs([u], [1.0], 0.37) - ref))) <= 1e-12 * scale, what
        if prog.corebox is not None:
            tiled += 1
            mask = _core_mask(prog)
            got = EmuKernel(plan, prog, tiled=True).rhs([u], [1.0], 0.37)
            assert float(np.max(np.abs(got[mask] - ref[mask]))) <= 1e-12 * scale, what
        plan.close()
    assert tiled >= 1


def random_interface_chain(rng):
    """Two to four 1-D domains joined end to end by interface conditions (written in either order), on a common uniform
    step or on random node vectors, upwind or WENO5 advection (linear or Burgers-like, either wind), optional diffusion
    (uniform grids), random Dirichlet / Neumann / no condition at the two outer ends."""
    t = sp.Symbol("t")
    nseg = int(rng.integers(2, 5))
    xs = sp.symbols("x1:%d" % (nseg + 1))
    us = [sp.Function("u%d" % (k + 1)) for k in range(nseg)]
    edges = np.concatenate([[0.0], np.cumsum(rng.uniform(0.3, 0.8, nseg))])
    uniform, weno = bool(rng.integers(2)), bool(rng.integers(2))
    diffusion = uniform and not weno and bool(rng.integers(2))
    dx = 0.02
    if uniform:
        edges = np.round(edges / dx) * dx
    grids = {}
    for k in range(nseg):
        if uniform:
            grids[xs[k]] = dx
        else:
            n = int(rng.integers(9, 20))
            g = np.sort(np.concatenate([[edges[k], edges[k + 1]], rng.uniform(edges[k], edges[k + 1], n - 2)]))
            if np.diff(g).min() < 1e-3 * (edges[k + 1] - edges[k]):
                g = np.linspace(edges[k], edges[k + 1], n)
            grids[xs[k]] = g
    eqs, bcs = [], []
    for k in range(nseg):
        U, Dx = us[k](t, xs[k]), Differential(xs[k])
        rhs = -float(rng.uniform(-1.5, 1.5)) * Dx(U) if rng.integers(2) else -U * Dx(U)
        if diffusion:
            rhs = rhs + 0.3 * (Dx ** 2)(U)
        if rng.integers(2):
            rhs = rhs + sp.sin(xs[k]) * U
        eqs.append(Eq(Differential(t)(U), rhs))
        bcs.append(Eq(us[k](0, xs[k]), sp.sin(2 * xs[k]) + 1.2))
    for k in range(nseg - 1):
        a, b = us[k](t, float(edges[k + 1])), us[k + 1](t, float(edges[k + 1]))
        bcs.append(Eq(a, b) if rng.integers(2) else Eq(b, a))
    for k, val in ((0, float(edges[0])), (nseg - 1, float(edges[-1]))):
        kind = str(rng.choice(["dirichlet", "neumann"] if weno else ["dirichlet", "neumann", "none"]))
        if kind == "dirichlet":
            bcs.append(Eq(us[k](t, val), sp.exp(-t)))
        elif kind == "neumann":
            bcs.append(Eq(Differential(xs[k])(us[k](t, val)), 0.1 * sp.cos(t)))
    dom = [Interval(t, 0.0, 1.0)] + [Interval(xs[k], float(edges[k]), float(edges[k + 1])) for k in range(nseg)]
    sys_ = PDESystem(eqs, bcs, dom, [t] + list(xs), [us[k](t, xs[k]) for k in range(nseg)])
    disc = MOLFiniteDifference(grids, t, advection_scheme=WENOScheme() if weno else UpwindScheme())
    return sys_, disc, f"{nseg} domains, uniform {uniform}, WENO {weno}, {[str(e) for e in eqs]}, {[str(b) for b in bcs[nseg:]]}"


@pytest.mark.parametrize("seed", range(2))
def test_stencil_program_matches_oracle_on_random_interface_chains(seed):
    from mol_b200.lowering import StencilLoweringError
    rng = np.random.default_rng(4000 + seed)
    done = 0
    while done < 8:
        sys_, disc, what = random_interface_chain(rng)
        try:
            prog = mol_b200.symbolic_discretize(sys_, disc)
        except StencilLoweringError as e:       # the reference's own ArgumentError (upwind_difference.jl:111-115)
            assert "extends past a non-interface boundary" in str(e), what
            continue
        done += 1
        orc = OracleProblem(sys_, disc)
        u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        ref = orc.rhs(u, 0.37)
        got = IRProgram(prog.text).rhs(u, 0.37)
        assert float(np.max(np.abs(ref - got))) <= 1e-12 * float(np.max(orc.rhs_termscale(u, 0.37))), what
