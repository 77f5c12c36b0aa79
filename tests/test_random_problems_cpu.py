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
