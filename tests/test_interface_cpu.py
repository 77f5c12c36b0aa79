"""Variables on different domains joined by interface boundary conditions (interface_boundary.jl:79-153)
and non-uniform periodic upwinding (upwind_difference.jl:85-129), without a GPU:

  * the oracle's chart coordinates against the reference's own known answers (test/Components/weno_interface_coords.jl),
  * the reference's solution-level acceptance tests on the oracle, at the reference's grid sizes and tolerances
    (test/Diffusion Test 14, test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl, test/Convection_WENO/
    MOL_1D_WENO_NU_Interface.jl incl. its order-of-convergence bar across the seam),
  * the reference's rejections (MOL_discretization.jl:55-95, interior_map.jl:33-52) raised by the lowering,
  * the product's side: chart layout of the stencil program, its Jacobian pattern, and (tests/cuda_emu) the generated
    unpack and JVP kernels; RHS parity of the generated kernels is in tests/test_ir_semantics_cpu.py /
    tests/test_generated_code_cpu.py (`iface_*` cases)."""
import numpy as np
import pytest
import sympy as sp

import mol_b200
from mol_b200 import capi
import problems as examples
from mol_b200.interface import Differential, Eq, Interval, MOLFiniteDifference, PDESystem, UpwindScheme, WENOScheme
from mol_b200.lowering import StencilLoweringError
from oracle.discretize import OracleProblem
from oracle.rk import solve_fixed, solve_tsit5

from cuda_emu import EmuKernel


def perturbed_grid(a, b, n, amp=0.004, seedmul=1.0):
    g = np.linspace(a, b, n)
    g[1:-1] += amp * np.sin(seedmul * np.arange(1, n - 1))
    return g


def test_oracle_bcoord_matches_reference_known_answers():
    # test/Components/weno_interface_coords.jl:118-166 ("Contiguous two-domain chart transition")
    g1, g2 = perturbed_grid(0.0, 0.5, 11, 0.002), perturbed_grid(0.5, 1.0, 16, 0.002, 2.0)
    sys_, disc = examples.advection_two_domains(x1grid=g1, x2grid=g2, scheme=WENOScheme())
    orc = OracleProblem(sys_, disc)
    N1 = len(g1)
    assert orc.bcoord(0, N1) == g1[N1 - 1]
    assert orc.bcoord(0, N1 + 1) == pytest.approx(g2[1], abs=1e-15)
    assert orc.bcoord(0, N1 + 2) == pytest.approx(g2[2], abs=1e-15)
    assert np.all(np.diff([orc.bcoord(0, N1 + i) for i in range(-2, 3)]) > 0)
    assert orc.bcoord(1, 2) == g2[1]
    assert orc.bcoord(1, 1) == pytest.approx(g1[N1 - 1], abs=1e-15)
    assert orc.bcoord(1, 0) == pytest.approx(g1[N1 - 2], abs=1e-15)
    assert orc.bcoord(1, -1) == pytest.approx(g1[N1 - 3], abs=1e-15)
    assert np.all(np.diff([orc.bcoord(1, 2 + i) for i in range(-2, 3)]) > 0)
    # the taps land on the neighbour's array (_wrapinterface, interface_boundary.jl:79-107)
    assert orc.wrap(0, N1 + 1) == (1, 2) and orc.wrap(1, 0) == (0, N1 - 1) and orc.wrap(1, 1) == (0, N1)
    # and the lowering's chart axis is the same chart
    prog = mol_b200.symbolic_discretize(sys_, disc)
    np.testing.assert_allclose(prog.axes[0].x, np.concatenate([g1, g2[1:]]), rtol=0, atol=1e-15)
    assert [(s["off"], s["n"]) for s in prog.segments] == [(0, 11), (10, 16)]
    assert prog.ilo == [[2], [12]] and prog.ihi == [[11], [25]]     # lower interface node clipped, upper kept


def _rel_l2(u, ref, x):
    w = np.append(np.diff(x), np.diff(x)[-1])
    return np.sqrt(np.sum(w * (u - ref) ** 2)) / np.sqrt(np.sum(w * ref ** 2))


def _solve_ssprk33(sys_, disc, dt, tmax):
    orc = OracleProblem(sys_, disc)
    nsteps = int(np.ceil(tmax / dt - 1e-9))
    ts, us = solve_fixed(orc.rhs, orc.u0, (0.0, tmax), tmax / nsteps, "ssprk33")
    return orc, orc.full_state(us[-1], tmax)


def test_oracle_diffusion_two_domains_reference_acceptance():
    # test/Diffusion/MOL_1D_Linear_Diffusion.jl:887-930 (Test 14): the joined solution decays to 0, atol 1e-3
    sys_, disc = examples.diffusion_two_domains()
    orc = OracleProblem(sys_, disc)
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 1.0), saveat=[1.0])
    c1, c2 = orc.full_state(us[-1], 1.0)
    solc = np.concatenate([c1, c2[1:]])
    assert len(solc) == 19 and np.all(np.abs(solc) <= 1e-3)
    assert c1[-1] == c2[0]


@pytest.mark.parametrize("order", [2, 4])
def test_oracle_two_independent_domains_reference_acceptance(order):
    # test/Diffusion/MOL_1D_Linear_Diffusion.jl:693-829 (Tests 12): u(t, x) and v(t, y) on their own domains in one system,
    # atol 0.01 against exp(-t) cos x / exp(-t) sin y at every saved time (30 points per domain here, 100 on the GPU)
    sys_, disc = examples.diffusion_two_independent_domains(l=30, approx_order=order)
    orc = OracleProblem(sys_, disc)
    saves = list(np.arange(0.0, 1.0 + 1e-9, 0.1))
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 1.0), saveat=saves)
    for t, u in zip(ts, us):
        U, V = orc.full_state(u, t)
        assert np.all(np.abs(U - np.exp(-t) * np.cos(orc.grid[0])) <= 0.01)
        assert np.all(np.abs(V - np.exp(-t) * np.sin(orc.grid[1])) <= 0.01)


def test_oracle_pde_with_ode_reference_acceptance():
    # test/Diffusion/MOL_1D_Linear_Diffusion.jl:830-885 (Test 13): exp(-t) sin x and exp(-t), atol 0.01
    sys_, disc = examples.diffusion_with_ode(l=30)
    orc = OracleProblem(sys_, disc)
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 1.0), saveat=list(np.arange(0.0, 1.0 + 1e-9, 0.1)))
    for t, u in zip(ts, us):
        U, V = orc.full_state(u, t)
        assert np.all(np.abs(U - np.exp(-t) * np.sin(orc.grid[0])) <= 0.01) and abs(V[0] - np.exp(-t)) <= 0.01
    prog = mol_b200.symbolic_discretize(sys_, disc)
    assert prog.segments[1]["n"] == 1 and prog.shapes[1] == (1,) and prog.nstate == orc.nstate


def test_field_driven_by_a_variable_of_t_alone():
    """One-way coupling: v(t) appears in the field's equation (token s:v reads v's single node); exact solution
    u = (1 + t) exp(-pi^2 t) ... is not needed: the oracle integrates the same system."""
    sys_, disc = examples.diffusion_driven_by_ode(l=20)
    prog = mol_b200.symbolic_discretize(sys_, disc)
    assert "s:1" in prog.text
    orc = OracleProblem(sys_, disc)
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 0.2), saveat=[0.2])
    U, V = orc.full_state(us[-1], 0.2)
    assert abs(V[0] - np.exp(-0.2)) <= 1e-3 and np.all(np.isfinite(U))
    plan = capi.Plan(prog.text, device=-1)
    colptr, rowval = plan.jac_sparsity()
    vcol = prog.offsets[1]                                              # the column of v: every field equation depends on it
    rows = set(rowval[colptr[vcol]:colptr[vcol + 1]])
    assert rows >= set(range(prog.offsets[0], prog.offsets[0] + prog.shapes[0][0]))
    plan.close()


def test_lowering_rejects_pointwise_coupling_across_domains():
    t, x1, x2 = sp.symbols("t x1 x2")
    u1, u2 = sp.Function("u1"), sp.Function("u2")
    Dt = Differential(t)
    eqs = [Eq(Dt(u1(t, x1)), (Differential(x1) ** 2)(u1(t, x1)) + u2(t, x2)), Eq(Dt(u2(t, x2)), (Differential(x2) ** 2)(u2(t, x2)))]
    bcs = [Eq(u1(0, x1), sp.sin(x1)), Eq(u2(0, x2), sp.sin(x2)), Eq(u1(t, 0.0), 0), Eq(u1(t, 1.0), 0), Eq(u2(t, 0.0), 0), Eq(u2(t, 2.0), 0)]
    dom = [Interval(t, 0.0, 1.0), Interval(x1, 0.0, 1.0), Interval(x2, 0.0, 2.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x1, x2], [u1(t, x1), u2(t, x2)])
    with pytest.raises(StencilLoweringError, match="another"):
        mol_b200.symbolic_discretize(sys_, MOLFiniteDifference({x1: 0.1, x2: 0.1}, t))


@pytest.mark.parametrize("case", ["positive_ratio400", "negative_symmetric"])
def test_oracle_interface_upwind_nonuniform_reference_acceptance(case):
    # test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl:497-546: rel L2 < 0.2 on both domains, continuity at the seam
    if case == "positive_ratio400":
        g1, g2 = examples.right_cluster_grid(0.0, 0.5, 51, 400.0), examples.one_sided_cluster_grid(0.5, 1.0, 51, 400.0)
        v, tmax = 1.0, 0.25
    else:
        g1, g2 = examples.symmetric_cluster_grid(0.0, 0.5, 71, 6.0), examples.symmetric_cluster_grid(0.5, 1.0, 71, 6.0)
        v, tmax = -1.2, 0.3
    assert np.diff(g1).max() / np.diff(g1).min() >= 50.0
    dt = min(0.25 * np.diff(g).min() / abs(v) for g in (g1, g2))
    sys_, disc = examples.advection_two_domains(x1grid=g1, x2grid=g2, v=v, tmax=tmax)
    orc, (u1, u2) = _solve_ssprk33(sys_, disc, dt, tmax)
    exact = lambda x: np.sin(2 * np.pi * (x - v * tmax))
    assert np.all(np.isfinite(u1)) and np.all(np.isfinite(u2))
    assert abs(u1[-1] - u2[0]) <= 0.05 + 0.05 * abs(u2[0])
    assert _rel_l2(u1, exact(g1), g1) < 0.2 and _rel_l2(u2, exact(g2), g2) < 0.2


def test_oracle_four_chained_interfaces_reference_acceptance():
    # same file :606-640: four chained non-uniform domains, rel L2 < 0.2 on each
    grids = [examples.right_cluster_grid(0.0, 0.25, 27, 60.0), examples.one_sided_cluster_grid(0.25, 0.5, 31, 40.0),
             examples.right_cluster_grid(0.5, 0.75, 25, 80.0), examples.one_sided_cluster_grid(0.75, 1.0, 29, 30.0)]
    dt = min(0.25 * np.diff(g).min() for g in grids)
    sys_, disc = examples.advection_chained_domains(grids=grids, v=1.0, tmax=0.1)
    orc, full = _solve_ssprk33(sys_, disc, dt, 0.1)
    for g, u in zip(grids, full):
        assert np.all(np.isfinite(u)) and _rel_l2(u, np.sin(2 * np.pi * (g - 0.1)), g) < 0.2
    for k in range(3):
        assert full[k][-1] == full[k + 1][0]


@pytest.mark.parametrize("case", ["positive_symmetric", "negative_chebyshev", "mass"])
def test_oracle_periodic_nonuniform_upwind_reference_acceptance(case):
    # same file :424-488: periodic advection on clustered grids, rel L2 < 0.2; Gaussian mass conserved to 5 %
    if case == "positive_symmetric":
        g, v, tmax, ic = examples.symmetric_cluster_grid(0.0, 1.0, 121, 5.0), 1.0, 0.4, None
    elif case == "negative_chebyshev":
        g, v, tmax, ic = examples.chebyshev_nodes(0.0, 1.0, 121), -2.0, 0.4, None
        assert np.diff(g).max() / np.diff(g).min() >= 20.0
    else:
        g, v, tmax = examples.symmetric_cluster_grid(0.0, 1.0, 141, 6.0), 0.6, 0.08
        ic = lambda xx: sp.exp(-((xx - 0.5) ** 2) / (2 * 0.04 ** 2))
    sys_, disc = examples.advection_periodic_speed(g, v=v, tmax=tmax, ic=ic)
    orc, (u,) = _solve_ssprk33(sys_, disc, 0.25 * np.diff(g).min() / abs(v), tmax)
    assert np.all(np.isfinite(u)) and u[0] == u[-1]
    if case == "mass":
        trapz = lambda y: float(np.sum((y[:-1] + y[1:]) * np.diff(g) / 2))
        m0 = trapz(np.exp(-((g - 0.5) ** 2) / (2 * 0.04 ** 2)))
        assert abs(trapz(u) - m0) <= 5e-2 * m0
    else:
        assert _rel_l2(u, np.sin(2 * np.pi * (g - v * tmax)), g) < 0.2


def test_oracle_weno_two_domain_interface_convergence_order():
    # test/Convection_WENO/MOL_1D_WENO_NU_Interface.jl:51-115: co-refined mismatched grids, Tsit5 at 1e-10;
    # errs[2] < 5e-3, every observed order > 3, the last > 3.3 -- no order loss across the seam
    errs = []
    for n1, n2 in ((41, 61), (81, 121), (161, 241)):
        sys_, disc = examples.weno_pulse_two_domains(n1, n2)
        orc = OracleProblem(sys_, disc)
        ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 0.5), abstol=1e-10, reltol=1e-10, saveat=[0.5])
        u1, u2 = orc.full_state(us[-1], 0.5)
        assert abs(u1[-1] - u2[0]) < 1e-8
        pulse = lambda x: np.exp(-((x - 0.5) - 0.7) ** 2 / (2 * 0.1 ** 2))
        errs.append(max(np.max(np.abs(u1 - pulse(orc.grid[0]))), np.max(np.abs(u2 - pulse(orc.grid[1])))))
    orders = [np.log2(errs[k] / errs[k + 1]) for k in range(2)]
    assert errs[1] < 5e-3 and all(o > 3.0 for o in orders) and orders[-1] > 3.3, (errs, orders)


def _two_domain_system(bcs_extra, eqs=None, scheme=None, g1=None, g2=None):
    t, x1, x2 = sp.symbols("t x1 x2")
    u1, u2 = sp.Function("u1"), sp.Function("u2")
    Dt, Dx1, Dx2 = Differential(t), Differential(x1), Differential(x2)
    eqs = eqs or [Eq(Dt(u1(t, x1)), -Dx1(u1(t, x1))), Eq(Dt(u2(t, x2)), -Dx2(u2(t, x2)))]
    if callable(eqs):
        eqs = eqs(t, x1, x2, u1, u2)
    bcs = [Eq(u1(0, x1), sp.sin(2 * sp.pi * x1)), Eq(u2(0, x2), sp.sin(2 * sp.pi * x2)), Eq(u1(t, 0.5), u2(t, 0.5))]
    bcs += bcs_extra(t, x1, x2, u1, u2)
    dom = [Interval(t, 0.0, 0.2), Interval(x1, 0.0, 0.5), Interval(x2, 0.5, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x1, x2], [u1(t, x1), u2(t, x2)])
    return sys_, MOLFiniteDifference({x1: g1, x2: g2}, t, advection_scheme=scheme or UpwindScheme())


DIRICHLET = lambda t, x1, x2, u1, u2: [Eq(u1(t, 0.0), 0.0), Eq(u2(t, 1.0), 0.0)]


@pytest.mark.parametrize("case", ["misaligned_grids", "scalar_vs_vector", "different_steps", "ring", "upwind_order_2",
                                  "higher_order_across_mismatched"])
def test_lowering_rejects_what_the_reference_rejects(case):
    """MOL_discretization.jl:55-95 (`_check_interface_boundarymap`), interior_map.jl:33-52, interface_boundary.jl:109-111;
    reference tests: test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl:253-375."""
    osc = examples.one_sided_cluster_grid
    if case == "misaligned_grids":
        args = dict(g1=osc(0.0, 0.5, 51, 200.0), g2=osc(0.55, 1.0, 51, 200.0))
    elif case == "scalar_vs_vector":
        args = dict(g1=0.01, g2=examples.symmetric_cluster_grid(0.5, 1.0, 41, 5.0))
    elif case == "different_steps":
        args = dict(g1=0.01, g2=0.02)
    elif case == "upwind_order_2":
        args = dict(g1=osc(0.0, 0.5, 21, 50.0), g2=osc(0.5, 1.0, 21, 50.0), scheme=UpwindScheme(2))
    elif case == "higher_order_across_mismatched":
        args = dict(g1=examples.sinus_stretched_grid(0.0, 0.5, 21, 0.01), g2=examples.sinus_stretched_grid(0.5, 1.0, 31, 0.012),
                    scheme=WENOScheme(),
                    eqs=lambda t, x1, x2, u1, u2: [
                        Eq(Differential(t)(u1(t, x1)), -Differential(x1)(u1(t, x1)) + 0.01 * (Differential(x1) ** 2)(u1(t, x1))),
                        Eq(Differential(t)(u2(t, x2)), -Differential(x2)(u2(t, x2)) + 0.01 * (Differential(x2) ** 2)(u2(t, x2)))])
    else:
        args = dict(g1=osc(0.0, 0.5, 31, 50.0), g2=osc(0.5, 1.0, 31, 50.0))
    extra = DIRICHLET if case != "ring" else (lambda t, x1, x2, u1, u2: [Eq(u2(t, 1.0), u1(t, 0.0))])
    sys_, disc = _two_domain_system(extra, **args)
    with pytest.raises(StencilLoweringError):
        mol_b200.symbolic_discretize(sys_, disc)


IFACE = {
    "iface_diffusion": lambda: examples.diffusion_two_domains(),
    "two_independent_domains": lambda: examples.diffusion_two_independent_domains(l=20),
    "pde_with_ode": lambda: examples.diffusion_with_ode(l=20),
    "pde_driven_by_ode": lambda: examples.diffusion_driven_by_ode(l=20),
    "iface_upwind_nu": lambda: examples.advection_two_domains(),
    "iface_upwind_chain4": lambda: examples.advection_chained_domains(),
    "iface_weno_nu_neg": lambda: examples.advection_two_domains(scheme=WENOScheme(), v=-1.0),
}


@pytest.mark.parametrize("name", sorted(IFACE))
def test_generated_unpack_kernel_on_the_chart_axis(name):
    """mol_unpack_full on the chart axis: each variable's own node range (what sol[u1(t, x1)] returns) equals the
    oracle's per-variable full-grid state, boundary and interface nodes included."""
    sys_, disc = IFACE[name]()
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    orc = OracleProblem(sys_, disc)
    emu = EmuKernel(plan, prog, unpack=True)
    u = orc.u0 + 0.05 * np.random.default_rng(3).standard_normal(orc.nstate)
    for t in (0.0, 0.37):
        got = emu.unpack(u, t).reshape(len(prog.ilo), -1)
        ref = orc.full_state(u, t)
        for v, seg in enumerate(prog.segments):
            np.testing.assert_allclose(got[v, seg["off"]:seg["off"] + seg["n"]], ref[v], rtol=0, atol=1e-13)
            np.testing.assert_array_equal(seg["x"], orc.grid[v])
    plan.close()


@pytest.mark.parametrize("name", sorted(IFACE))
def test_generated_jvp_and_jacobian_pattern_across_interfaces(name):
    sys_, disc = IFACE[name]()
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    orc = OracleProblem(sys_, disc)
    n = orc.nstate
    rng = np.random.default_rng(21)
    u, v, t = orc.u0 + 0.05 * rng.standard_normal(n), rng.standard_normal(n), 0.37
    got = EmuKernel(plan, prog, jvp=True).jvp(u, v, t)
    if "weno" in name:
        want = (orc.rhs(u + 1e-6 * v, t) - orc.rhs(u - 1e-6 * v, t)) / 2e-6
        assert np.max(np.abs(got - want)) <= 2e-6 * max(1.0, float(np.max(np.abs(want))))
    else:                                                                   # affine in u: J v = f(v) - f(0), exactly
        want = orc.rhs(v, t) - orc.rhs(np.zeros(n), t)
        assert np.max(np.abs(got - want)) <= 1e-12 * float(np.max(orc.rhs_termscale(v, t)))
    colptr, rowval = plan.jac_sparsity()
    pattern = np.zeros((n, n), dtype=bool)
    for j in range(n):
        pattern[rowval[colptr[j]:colptr[j + 1]], j] = True
    f0 = orc.rhs(u, t)
    numeric = np.zeros((n, n), dtype=bool)
    for j in range(n):
        up = u.copy(); up[j] += 1e-6
        numeric[:, j] = np.abs(orc.rhs(up, t) - f0) > 1e-9 * 1e-6 * max(1.0, float(np.max(np.abs(f0))))
    assert not (numeric & ~pattern).any()
    # the coupling across the seam is in the pattern: some equation of one variable reads an unknown of the other
    o1 = prog.offsets[1]
    assert (pattern[:o1, o1:].any() or pattern[o1:, :o1].any()) == (name not in ("two_independent_domains", "pde_with_ode"))  # (pde_driven_by_ode: coupled through s:v)
    plan.close()
