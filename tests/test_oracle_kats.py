"""Pins the oracle against the reference's own known-answer tests and literal artifacts (CPU only).

Every expected value below is copied from a reference TEST or DOC (cited per test), never from the
oracle itself."""
import json
import os

import numpy as np
import pytest
import sympy as sp

from oracle import operators as ops
from oracle import weno as wk
from oracle.discretize import OracleProblem
from oracle.fornberg import calculate_weights

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fornberg_kats():
    # test/Components/MOLfornberg_weights.jl:8-30 (exact ==)
    assert np.array_equal(calculate_weights(2, 0.0, [-1, 0, 1.0]), [1, -2, 1])
    assert np.array_equal(calculate_weights(1, 0.0, [-1.0, 1.0]), [-0.5, 0.5])
    assert np.array_equal(calculate_weights(1, 0.0, [0, 1]), [-1, 1])
    assert np.array_equal(calculate_weights(1, 1.0, [0, 1]), [-1, 1])
    assert np.array_equal(calculate_weights(3, 0.0, [0, 1, 2, 3, 4, 5]),
                          [-17 / 4, 71 / 4, -59 / 2, 49 / 2, -41 / 4, 7 / 4])


def test_centered_tables():
    # test/shared/finite_diff_schemes.jl:23-30 (dx = 1)
    exp2 = ([-0.5, 0, 0.5], [1.0, -2.0, 1.0], [-1 / 2, 1.0, 0.0, -1.0, 1 / 2])
    exp4 = ([1 / 12, -2 / 3, 0, 2 / 3, -1 / 12], [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12],
            [1 / 8, -1.0, 13 / 8, 0.0, -13 / 8, 1.0, -1 / 8])
    for p, exp in ((2, exp2), (4, exp4)):
        for d in (1, 2, 3):
            np.testing.assert_allclose(ops.centered(d, p, 1.0).stencil_coefs, exp[d - 1], rtol=0, atol=1e-13)
    # SURVEY App. A.3 spot rows (one-sided rows used by Neumann/Robin edges and order-4 frames)
    np.testing.assert_allclose(ops.centered(1, 2, 1.0).low_boundary_coefs[0], [-1.5, 2.0, -0.5], atol=1e-14)
    np.testing.assert_allclose(ops.centered(2, 2, 1.0).low_boundary_coefs[0], [2, -5, 4, -1], atol=1e-13)
    np.testing.assert_allclose(ops.centered(2, 4, 1.0).low_boundary_coefs[1],
                               [5 / 6, -5 / 4, -1 / 3, 7 / 6, -1 / 2, 1 / 12], atol=1e-13)
    # half-point interpolation (order max(4,p)) and extrapolation rows (App. A.5)
    np.testing.assert_allclose(ops.half_centered(0, 4, 1.0).stencil_coefs, [-1 / 16, 9 / 16, 9 / 16, -1 / 16], atol=1e-15)
    np.testing.assert_allclose(ops.extrapolator(6, 1.0).low_boundary_coefs[0], [0, 5, -10, 10, -5, 1], atol=1e-12)


def test_periodic_wrap():
    # test/Components/utils_test.jl:129-138: _wrapperiodic(I, N, j, l)
    P = OracleProblem.__new__(OracleProblem)
    assert P._wrap(5, 4) == 2
    assert P._wrap(1, 4) == 4
    assert P._wrap(-1, 4) == 2


def test_literal_generated_rhs_brusselator():
    # docs/src/generated/bruss_code.md:82-113 evaluated on seeded inputs (tests/golden/make_bruss_golden.py)
    import problems as ex
    G = json.load(open(os.path.join(GOLD, "bruss_code_n4.json")))
    sys_, disc = ex.brusselator_2d(4)
    P = OracleProblem(sys_, disc)
    assert P.nstate == 32
    for case in G["cases"]:
        du = P.rhs(np.array(case["u"]), 0.0)
        ref = np.array(case["du"])
        assert np.max(np.abs(du - ref)) / np.max(np.abs(ref)) <= 1e-14, case["seed"]
    # self-check stated in the doc text: u = v = 1 -> -2.4 / +2.4
    du = P.rhs(np.ones(32), 0.0)
    np.testing.assert_allclose(du[:16], -2.4, atol=1e-11)
    np.testing.assert_allclose(du[16:], 2.4, atol=1e-11)


def test_weno_uniform_kernel():
    # inputs: test/Components/weno_dispatch.jl:8, benchmark/weno/suite.jl:13; linear data -> exact slope
    assert abs(wk.weno_f_uniform([1.0, 2.0, 3.0, 4.0, 5.0], 1e-6, 0.1) - 10.0) < 1e-13
    v = wk.weno_f_uniform([1.3, 2.1, 1.7, 0.4, 0.9], 1e-6, 0.1)
    assert np.isfinite(v)


def test_weno_nonuniform_core_properties():
    # test/Components/weno_nonuniform_core.jl:53-75 (polynomial exactness, degree <= 2)
    xs = np.array([0.0, 0.6, 1.4, 2.1, 3.3])
    xc = xs[2]
    f = lambda u: wk.weno_f_nonuniform_core(list(u), 1e-6, list(xs), 3)
    assert abs(f(np.full(5, 1.7)) - 0.0) <= 1e-14
    assert abs(f(1.7 + 0.9 * xs) - 0.9) <= 1e-13
    assert abs(f(1.7 + 0.9 * xs - 0.4 * xs ** 2) - (0.9 - 0.8 * xc)) <= 1e-12
    assert abs(f(xs ** 3) - 3 * xc ** 2) > 1e-6
    # order of convergence (MMS), :95-118
    o = np.array([-2.0, -1.13, 0.08, 0.91, 2.0])
    errs = []
    for h in (0.2, 0.1, 0.05, 0.025, 0.0125):
        x = 1.0 + h * o
        errs.append(abs(wk.weno_f_nonuniform_core(list(np.sin(1.3 * x) + 0.5 * x), 1e-6, list(x), 3)
                        - (1.3 * np.cos(1.3 * x[2]) + 0.5)))
    orders = [np.log2(errs[k] / errs[k + 1]) for k in range(4)]
    assert orders[-1] > 3.85 and orders[-2] > 3.7 and all(o_ > 3.0 for o_ in orders)


@pytest.mark.parametrize("xs", [[0.0, 0.3, 0.9, 1.7, 2.2], [-0.3, 0.4, 0.55, 1.9, 2.4]])
@pytest.mark.parametrize("T", [1, 2, 4, 5])
def test_weno_nonuniform_boundary_targets(xs, T):
    # test/Components/weno_nonuniform_boundary.jl:87-99
    xs = np.array(xs)
    xt = xs[T - 1]
    f = lambda u: wk.weno_f_nonuniform_core(list(u), 1e-6, list(xs), T)
    assert abs(f(np.full(5, 1.7))) <= 1e-13
    assert abs(f(1.7 + 0.9 * xs) - 0.9) <= 1e-13
    assert abs(f(1.7 + 0.9 * xs - 0.4 * xs ** 2) - (0.9 - 0.8 * xt)) <= 1e-12


def test_heat_dirichlet_matches_analytic():
    # test/Diffusion/MOL_1D_Linear_Diffusion.jl:26-85 acceptance: |u - e^-t cos x| <= 0.01
    import problems as ex
    from oracle.rk import solve_tsit5
    sys_, disc = ex.heat_1d_dirichlet(dx=0.05)
    P = OracleProblem(sys_, disc)
    ts, us, stats = solve_tsit5(P.rhs, P.u0, (0.0, 1.0), saveat=[0.5, 1.0])
    x = P.grid[0][1:-1]
    for t, u in zip(ts, us):
        assert np.max(np.abs(u - np.exp(-t) * np.cos(x))) <= 0.01


def _brusselator_2d_loop(u, N, t):
    """The reference's INDEPENDENT hand-written RHS (test/Brusselator/brusselator_eq.jl:79-105, `brusselator_2d_loop`),
    vectorised: grid xyd = 0:dx:(N-1)dx, periodic neighbours, p = (A, B, alpha, dx) = (3.4, 1.0, 10.0, 1/N)."""
    A, B, alpha = 3.4, 1.0, 10.0 * N * N
    xy = np.arange(N) / N
    X, Y = np.meshgrid(xy, xy, indexing="ij")
    f = (((X - 0.3) ** 2 + (Y - 0.6) ** 2) <= 0.1 ** 2) * (t >= 1.1) * 5.0
    lap = lambda w: np.roll(w, 1, 0) + np.roll(w, -1, 0) + np.roll(w, 1, 1) + np.roll(w, -1, 1) - 4 * w
    U, V = u[..., 0], u[..., 1]
    du = np.empty_like(u)
    du[..., 0] = alpha * lap(U) + B + U ** 2 * V - (A + 1) * U + f
    du[..., 1] = alpha * lap(V) + A * U - U ** 2 * V
    return du


@pytest.mark.parametrize("t", [0.0, 2.0])
def test_brusselator_independent_loop_rhs(t):
    """MOL's unknowns sit at x = dx..1 (nodes 2..N+1), the loop's at 0..(N-1)dx: MOL node i in 2..N is loop index i,
    MOL node N+1 (x = 1, the periodic image of x = 0) is loop index 1 (the reference test compares the two
    solutions as solu[2:end, 2:end] vs msol, brusselator_eq.jl:124-131).  On that mapping the Laplacian and the reaction
    terms agree to rounding; the forcing disc is sampled at the same physical points except along x = 1 / y = 1 (MOL)
    vs x = 0 / y = 0 (loop), where the disc centred at (0.3, 0.6) with radius 0.1 vanishes on both -- so the whole RHS
    agrees, forcing on (t = 2) or off."""
    import problems as ex
    N = 32
    orc = OracleProblem(*ex.brusselator_2d(N))
    rng = np.random.default_rng(5)
    state = rng.uniform(0.0, 3.0, orc.nstate)
    # MOL state: u block then v block, x fastest, nodes 2..N+1  ->  loop array [i, j, species], loop index = node mod N
    mol = state.reshape(2, N, N).transpose(2, 1, 0)            # [ix, iy, species] with ix = node - 2
    loop_u = np.roll(mol, shift=(1, 1), axis=(0, 1))            # node N+1 (ix = N-1) -> loop index 0, node 2 -> index 1
    du_loop = _brusselator_2d_loop(loop_u, N, t)
    du_mol = np.roll(du_loop, shift=(-1, -1), axis=(0, 1)).transpose(2, 1, 0).reshape(-1)
    ref = orc.rhs(state, t)
    assert np.max(np.abs(ref - du_mol)) <= 1e-12 * np.max(np.abs(ref))
    if t >= 1.1:
        assert np.count_nonzero(_brusselator_2d_loop(np.zeros((N, N, 2)), N, t)[..., 0] - 1.0) > 10     # the forcing is on


def test_mixed_derivative_is_consistent_with_the_analytic_one():
    """The reference does not pin mixed derivatives numerically (test/Mixed_Derivatives only smoke-tests them), so the
    oracle's restatement of mixed_central_difference (2nd_order_mixed_deriv.jl:5-22) is checked against calculus:
    u = sin(x + y) + 1 gives u_xx + u_yy + k u_xy = -(2 + k) sin(x + y), second-order accurate away from the corner
    nodes (which the reference reads as 0, generate_bc_eqs.jl:396-416)."""
    import problems as ex
    errs = []
    for n in (20, 40):
        orc = OracleProblem(*ex.anisotropic_diffusion_2d(n, n, kxy=0.5))
        X, Y = np.meshgrid(orc.grid[0], orc.grid[1], indexing="ij")
        sl = orc._islice(0)
        u = (np.sin(X + Y) + 1)[sl].reshape(-1, order="F")
        du = orc.rhs(u, 0.0).reshape(orc.ishape[0], order="F")
        exact = (-2.5 * np.sin(X + Y))[sl]
        inner = (slice(1, -1), slice(1, -1))                 # nodes whose mixed stencil does not touch a corner node
        errs.append(np.max(np.abs(du[inner] - exact[inner])))
        corner_err = abs(du[0, 0] - exact[0, 0])
        assert corner_err > 10 * errs[-1]                     # the corner node is read as 0, not as the boundary datum
    assert errs[0] < 2e-3 and 3.5 < errs[0] / errs[1] < 4.5, errs


def _dirichlet_advection(grid_or_dx, scheme):
    import mol_b200
    from mol_b200.interface import Differential, Eq, Interval, MOLFiniteDifference, PDESystem
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    eq = Eq(Differential(t)(u(t, x)), -Differential(x)(u(t, x)))
    bcs = [Eq(u(0, x), sp.sin(sp.pi * x)), Eq(u(t, 0.0), 0.0), Eq(u(t, 1.0), 0.0)]
    sys_ = PDESystem([eq], bcs, [Interval(t, 0.0, 1.0), Interval(x, 0.0, 1.0)], [t, x], [u(t, x)])
    return sys_, MOLFiniteDifference({x: grid_or_dx}, t, advection_scheme=scheme)


def test_interior_map_extents_uniform_vs_nonuniform_weno():
    """test/Components/weno_boundary_integration.jl:54-86: with WENOScheme the stencil extents are ([2], [2]) on a uniform
    grid and ([0], [0]) on a node vector; 21 nodes: the interior is 2..20 (non-uniform) / 3..19 (uniform).  Checked on the
    oracle and on the lowering."""
    import mol_b200
    from mol_b200 import WENOScheme
    g = np.linspace(0.0, 1.0, 21)
    g[1:-1] += 0.004 * np.sin(np.arange(1, 20))
    for spec, ext, ilo, ihi in ((g, ([0], [0]), 2, 20), (1 / 20, ([2], [2]), 3, 19)):
        sys_, disc = _dirichlet_advection(spec, WENOScheme())
        orc = OracleProblem(sys_, disc)
        assert orc.ext[0] == ext and orc.ilo == [[ilo]] and orc.ihi == [[ihi]]
        prog = mol_b200.symbolic_discretize(sys_, disc)
        assert prog.ilo == [[ilo]] and prog.ihi == [[ihi]]


def test_grid_construction_known_answers():
    """test/Components/DiscreteSpace.jl:42-46,89-93: centre-aligned grid x_min:dx:x_max, edge-aligned grid
    (x_min - dx/2):dx:(x_max + dx/2), for dx = 0.1 and dy = 0.2 on [0, 2] (Julia range values = exact rational arithmetic
    rounded once per node)."""
    from fractions import Fraction
    from oracle.discretize import make_grid
    from mol_b200.lowering import Axis
    for dx in (0.1, 0.2):
        step = Fraction(dx).limit_denominator(1000)
        n = int(2 / dx + 0.5) + 1
        centre = np.array([float(k * step) for k in range(n)])
        edge = np.array([float(-step / 2 + k * step) for k in range(n + 1)])
        g, d = make_grid(0.0, 2.0, dx)
        assert d == dx and np.array_equal(g, centre)
        g, d = make_grid(0.0, 2.0, dx, edge=True)
        assert np.array_equal(g, edge)
        assert np.array_equal(Axis(sp.Symbol("x"), 0.0, 2.0, dx).x, centre)
        assert np.array_equal(Axis(sp.Symbol("x"), 0.0, 2.0, dx, edge=True).x, edge)
