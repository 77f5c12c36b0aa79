"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the oracle on identical seeded inputs.

Bar (SURVEY §7 hard part 1): FP64, per RHS evaluation
  (a) max|du_gpu - du_oracle| <= 1e-12 * max|du_oracle|            on generic (seeded random / perturbed) states, and
  (b) max|du_gpu - du_oracle| <= 1e-13 * max_i sum|terms_i|        on every state, where sum|terms_i| is the oracle's
      magnitude evaluation (OracleProblem.rhs_termscale): the scale rounding is proportional to.
(b) is the meaningful bound on smooth states, where du is O(1) but the stencil terms are O(1/dx^2)
and cancel (e.g. u = cos x, dx = 0.01: terms ~ 4e4, du ~ 1, so 1 ulp of a term is 1e-12 of du)."""
import json
import os

import numpy as np
import pytest

import mol_b200
from mol_b200 import capi
import problems as examples

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLD = os.path.join(os.path.dirname(__file__), "golden")


TOL_TERMS = 1e-13


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def terms_tol(prob):
    """Bar (b).  Non-uniform WENO5 is evaluated in a different, better conditioned arithmetic than the reference's
    (tests/test_weno_nu_accuracy_cpu.py): the two agree to the REFERENCE's rounding, a few eps * |x| / h on the grid."""
    P = prob.program
    if "\nwtab " not in P.text:
        return TOL_TERMS
    cond = max((np.max(np.abs(ax.x)) / np.min(np.diff(ax.x)) for ax in P.axes if not ax.uniform), default=0.0)
    return max(TOL_TERMS, 4 * np.finfo(float).eps * cond)


def check_rhs(prob, orc, u, t, mode, generic_state, tag=""):
    got = gpu_rhs(prob, u, t, mode)
    ref = orc.rhs(u, t)
    scale = float(np.max(orc.rhs_termscale(u, t)))
    err = float(np.max(np.abs(got - ref)))
    assert err <= terms_tol(prob) * scale, (tag, mode, t, "vs term scale", err / scale)
    if generic_state:
        assert err <= TOL * np.max(np.abs(ref)), (tag, mode, t, "vs max|du|", err / np.max(np.abs(ref)))


def gpu_rhs(prob, u, t, mode=capi.KERNEL_AUTO):
    prob.plan.set_option("kernel", mode)
    return prob.rhs_host(u, t)


def oracle_for(sys_, disc):
    from oracle.discretize import OracleProblem
    return OracleProblem(sys_, disc)


def test_brusselator_literal_reference_rhs():
    """The reference's own generated code (docs/src/generated/bruss_code.md) on seeded inputs."""
    G = json.load(open(os.path.join(GOLD, "bruss_code_n4.json")))
    prob = mol_b200.discretize(*examples.brusselator_2d(4))
    for case in G["cases"]:
        ref = np.array(case["du"])
        for mode in (capi.KERNEL_AUTO, capi.KERNEL_GENERIC):
            assert relmax(gpu_rhs(prob, np.array(case["u"]), 0.0, mode), ref) <= TOL


@pytest.mark.parametrize("N", [32, 130, 512])
def test_brusselator_vs_oracle(N):
    sys_, disc = examples.brusselator_2d(N)
    prob = mol_b200.discretize(sys_, disc)
    orc = oracle_for(sys_, disc)
    assert np.array_equal(prob.u0, orc.u0)
    rng = np.random.default_rng(0)
    for generic, u in ((False, orc.u0), (True, rng.uniform(0.0, 3.0, orc.nstate))):
        for t in (0.0, 2.0):
            check_rhs(prob, orc, u, t, capi.KERNEL_AUTO, generic, N)
            if N <= 130:
                check_rhs(prob, orc, u, t, capi.KERNEL_GENERIC, generic, N)


def test_brusselator_4096_properties():
    """Full BASELINE size: size-independent properties + the C restatement of the generated RHS."""
    import torch
    from oracle import cref
    N = 4096
    sys_, disc = examples.brusselator_2d(N)
    prob = mol_b200.discretize(sys_, disc)
    n = prob.plan.state_len
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    u = rng.uniform(0.0, 3.0, n)
    ud = torch.from_numpy(u).to(dev)
    du = torch.empty_like(ud)
    # (1) value-identical C restatement of the reference's per-point expression, t = 0 and t = 2
    xg, yg = prob.program.axes[0].x, prob.program.axes[1].x
    for t in (0.0, 2.0):
        prob.f(du, ud, None, t)
        ref = cref.bruss_rhs(u, xg, yg, N, t, nthreads=cref.lib().bruss_ref_max_threads())
        assert relmax(du.cpu().numpy(), ref) <= TOL
    # (2) periodic translation invariance (t = 0, no forcing): f(shift u) == shift f(u), bit for bit
    U = ud.view(2, N, N)
    sh = torch.roll(U, shifts=(17, 5), dims=(1, 2)).contiguous().view(-1)
    du2 = torch.empty_like(ud)
    prob.f(du, ud, None, 0.0)
    prob.f(du2, sh, None, 0.0)
    assert torch.equal(torch.roll(du.view(2, N, N), shifts=(17, 5), dims=(1, 2)).contiguous().view(-1), du2)
    # (3) constant state: the Laplacian vanishes identically
    c = torch.cat([torch.full((N * N,), 1.5, dtype=torch.float64), torch.full((N * N,), 0.7, dtype=torch.float64)]).to(dev)
    prob.f(du, c, None, 0.0)
    out = du.cpu().numpy()
    np.testing.assert_allclose(out[:N * N], 1.0 + 1.5 ** 2 * 0.7 - 4.4 * 1.5, rtol=0, atol=1e-6)
    np.testing.assert_allclose(out[N * N:], 3.4 * 1.5 - 1.5 ** 2 * 0.7, rtol=0, atol=1e-6)


def _edge(sys_, disc):
    """The same problem on an edge-aligned grid (grid_align = edge_align, src/interface/grid_types.jl:1-33)."""
    return sys_, mol_b200.MOLFiniteDifference(disc.dxs, disc.time, approx_order=disc.approx_order,
                                              advection_scheme=disc.advection_scheme, grid_align=mol_b200.edge_align)


CASES = {
    # edge-aligned grids: boundary values / derivatives through half-offset interpolation rows (generate_bc_eqs.jl:79-161)
    "edge_heat_neumann": lambda: _edge(*examples.heat_1d_neumann(dx=0.05)),
    "edge_heat_robin_o4": lambda: _edge(*examples.heat_1d_robin_order4(dx=0.05)),
    "edge_burgers2d": lambda: _edge(*examples.burgers_2d(nx=40, ny=36)),
    "heat_dirichlet": lambda: examples.heat_1d_dirichlet(dx=0.01),
    "heat_dirichlet_o4": lambda: examples.heat_1d_dirichlet(dx=0.02, approx_order=4),
    "heat_neumann": lambda: examples.heat_1d_neumann(dx=0.05),
    "heat_robin": lambda: examples.heat_1d_robin(dx=0.05),
    "burgers_upwind": lambda: examples.burgers_1d(dx=0.02),
    "burgers_upwind_nu": lambda: examples.burgers_1d(
        grid=np.sort(np.concatenate([[0.0, 1.0], np.random.default_rng(0).uniform(0.02, 0.98, 40)]))),
    "burgers_weno": lambda: examples.burgers_1d(dx=0.02, scheme=mol_b200.WENOScheme()),
    "advection_weno_periodic": lambda: examples.advection_1d_periodic(dx=0.02, scheme=mol_b200.WENOScheme()),
    "advection_upwind_periodic": lambda: examples.advection_1d_periodic(dx=0.02),
    "nonlinear_diffusion": lambda: examples.nonlinear_diffusion_1d(dx=0.02),
    "spherical": lambda: examples.spherical_diffusion_1d(dr=0.05),
    "burgers2d": lambda: examples.burgers_2d(nx=40, ny=36),
    "burgers2d_nu": lambda: examples.burgers_2d(
        grid_x=0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, 41)) / np.tanh(2.0)),
        grid_y=np.linspace(0, 1, 37) ** 1.3),
    # benchmark/weno/grids.jl grid kinds through the non-uniform WENO path (periodic wrap of chart coordinates)
    "advection_weno_stretched": lambda: examples.advection_1d_periodic(dx=examples.stretched_grid(0, 2, 64), scheme=mol_b200.WENOScheme()),
    "advection_weno_uniform_vector": lambda: examples.advection_1d_periodic(dx=np.linspace(0, 2, 64), scheme=mol_b200.WENOScheme()),
    "weno_burgers_periodic": lambda: examples.weno_burgers_periodic(dx=2.0 / 63),
    "weno_burgers_stretched": lambda: examples.weno_burgers_periodic(dx=examples.stretched_grid(0, 2, 64)),
    "burgers_weno_nu_dirichlet": lambda: examples.burgers_1d(grid=examples.stretched_grid(0, 1, 41, 0.03), scheme=mol_b200.WENOScheme()),
    "advection2d_weno": lambda: examples.advection_2d_periodic(40, scheme=mol_b200.WENOScheme()),
    # non-uniform WENO5 through the TILED kernel (per-interval geometry arrays): several 1-D tiles with the periodic
    # seam, walls (records on the frame rows), 2-D stretched in both directions
    "advection_weno_stretched_5000": lambda: examples.advection_1d_periodic(dx=examples.stretched_grid(0, 2, 5000), scheme=mol_b200.WENOScheme()),
    "burgers_weno_nu_dirichlet_4097": lambda: examples.burgers_1d(grid=examples.stretched_grid(0, 1, 4097, 0.03), scheme=mol_b200.WENOScheme()),
    "advection2d_weno_nu": lambda: examples.advection_2d_periodic(scheme=mol_b200.WENOScheme(), grid_x=examples.stretched_grid(0, 2, 131),
                                                                  grid_y=examples.sinus_stretched_grid(0, 2, 91, 0.1)),
    "advection2d_upwind_o4_diffusion": lambda: examples.advection_2d_periodic(40, nu=0.01, approx_order=4),
    "brusselator_o4": lambda: examples.brusselator_2d(40, approx_order=4),
    "fisher3d_periodic": lambda: examples.diffusion_reaction_3d(n=20, periodic=True),
    "fisher3d_dirichlet_z": lambda: examples.diffusion_reaction_3d(n=20, periodic=False),
    # several xy tiles and two z chunks of the z-marching kernel (64 x 16 tiles, 32 planes per work item)
    "fisher3d_periodic_multitile": lambda: examples.diffusion_reaction_3d(n=72, periodic=True, nz=40),
    "fisher3d_dirichlet_z_multitile": lambda: examples.diffusion_reaction_3d(n=72, periodic=False, nz=40),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_scheme_parity(name):
    """Every scheme family / boundary rule of SURVEY §8a against the oracle, both kernels."""
    sys_, disc = CASES[name]()
    prob = mol_b200.discretize(sys_, disc)
    orc = oracle_for(sys_, disc)
    assert prob.plan.state_len == orc.nstate
    # ICs are evaluated by two independent evaluators (lambdify vs the oracle's walker): product order may differ by 1 ulp
    np.testing.assert_allclose(prob.u0, orc.u0, rtol=1e-15, atol=1e-16)
    rng = np.random.default_rng(7)
    states = [orc.u0, orc.u0 + 0.05 * rng.standard_normal(orc.nstate)]
    if name.startswith("nonlinear") or name.startswith("spherical"):
        states[1] = np.abs(states[1]) + 0.1
    for k, u in enumerate(states):
        for t in (0.0, 0.37):
            for mode in (capi.KERNEL_AUTO, capi.KERNEL_GENERIC):
                check_rhs(prob, orc, u, t, mode, k == 1, name)


def test_tsit5_heat_matches_analytic_and_oracle():
    """Config 1: 1-D heat, Dirichlet, 101 points, Tsit5 (docs/src/tutorials/heat.md:19-41)."""
    from oracle.rk import solve_tsit5
    sys_, disc = examples.heat_1d_dirichlet(dx=0.01)
    prob = mol_b200.discretize(sys_, disc)
    sol = mol_b200.solve(prob, mol_b200.Tsit5(), saveat=0.2)
    assert sol.retcode == "Success"
    x = prob.program.axes[0].x[1:-1]
    for t, u in zip(sol.t, sol.u):
        assert np.max(np.abs(u - np.exp(-t) * np.cos(x))) <= 0.01      # reference acceptance (test/Diffusion ...:82)
    orc = oracle_for(sys_, disc)
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 1.0), saveat=[1.0])
    # both integrate to abstol=1e-6 / reltol=1e-3
    assert np.max(np.abs(sol.u[-1] - us[-1])) <= 1e-6 + 1e-3 * np.max(np.abs(us[-1]))


def test_tsit5_tight_tolerance_vs_oracle_brusselator():
    from oracle.rk import solve_tsit5
    sys_, disc = examples.brusselator_2d(16, tmax=0.2)
    prob = mol_b200.discretize(sys_, disc)
    sol = mol_b200.solve(prob, mol_b200.Tsit5(), abstol=1e-10, reltol=1e-10)
    orc = oracle_for(sys_, disc)
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 0.2), abstol=1e-10, reltol=1e-10)
    np.testing.assert_allclose(sol.u[-1], us[-1], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("alg", ["euler", "ssprk33", "rk4", "tsit5"])
def test_fixed_step_methods_vs_oracle(alg):
    """Fixed-dt SSPRK33 / Euler as in benchmark/weno/suite.jl:50-54, test/Convection/...:45."""
    from oracle.rk import solve_fixed
    sys_, disc = examples.advection_1d_periodic(dx=0.02, scheme=mol_b200.WENOScheme(), tmax=0.2)
    prob = mol_b200.discretize(sys_, disc)
    A = {"euler": mol_b200.Euler(), "ssprk33": mol_b200.SSPRK33(), "rk4": mol_b200.RK4(), "tsit5": mol_b200.Tsit5()}[alg]
    dt = 0.4 * 0.02
    sol = mol_b200.solve(prob, A, dt=dt, adaptive=False)
    orc = oracle_for(sys_, disc)
    ts, us = solve_fixed(orc.rhs, orc.u0, (0.0, 0.2), dt, alg)
    np.testing.assert_allclose(sol.u[-1], us[-1], rtol=0, atol=1e-11)


def test_saveat_is_dense_output_and_does_not_change_the_step_sequence():
    """OrdinaryDiffEq's saveat: states at the save points come from the method's interpolant inside the covering step
    (Tsit5: free 4th-order interpolant rebuilt from u, u+, k1..k5, k7 -- k6 is never stored), only t1 is a stop time."""
    from oracle.rk import solve_tsit5
    sys_, disc = examples.heat_1d_dirichlet(dx=0.01)
    prob = mol_b200.discretize(sys_, disc)
    orc = oracle_for(sys_, disc)
    plain = mol_b200.solve(prob, mol_b200.Tsit5(), abstol=1e-8, reltol=1e-8)
    sv = np.array([0.0, 0.0137, 0.2, 0.2000001, 0.731, 1.0])
    sol = mol_b200.solve(prob, mol_b200.Tsit5(), abstol=1e-8, reltol=1e-8, saveat=sv)
    assert sol.retcode == "Success" and sol.stats["naccept"] == plain.stats["naccept"] and sol.stats["nf"] == plain.stats["nf"]
    ts, us, st = solve_tsit5(orc.rhs, orc.u0, (0.0, 1.0), abstol=1e-8, reltol=1e-8, saveat=sv)
    assert st["naccept"] == sol.stats["naccept"] and st["nreject"] == sol.stats["nreject"]
    x = prob.program.axes[0].x[1:-1]
    for k in range(len(sv)):
        np.testing.assert_allclose(sol.u[k], us[k], rtol=0, atol=1e-10)
        assert np.max(np.abs(sol.u[k] - np.exp(-sv[k]) * np.cos(x))) <= 2e-5        # discretisation error of dx = 0.01
    np.testing.assert_allclose(sol.u[-1], plain.u[-1], rtol=0, atol=1e-13)


@pytest.mark.parametrize("alg", ["euler", "ssprk33", "rk4", "tsit5"])
def test_fixed_step_saveat_between_steps_and_last_step_clipped_to_t1(alg):
    """dt does not divide the span and the save points fall between steps: the last step is shortened to land on t1 and
    the saved states are interpolated (Hermite / Tsit5 interpolant) -- never uninitialised memory."""
    from oracle.rk import solve_fixed
    sys_, disc = examples.advection_1d_periodic(dx=0.02, scheme=mol_b200.WENOScheme(), tmax=0.1)
    prob = mol_b200.discretize(sys_, disc)
    A = {"euler": mol_b200.Euler(), "ssprk33": mol_b200.SSPRK33(), "rk4": mol_b200.RK4(), "tsit5": mol_b200.Tsit5()}[alg]
    dt = 0.0071
    sv = [0.0, 0.01, 0.05, 0.0999, 0.1]
    sol = mol_b200.solve(prob, A, dt=dt, adaptive=False, saveat=sv)
    orc = oracle_for(sys_, disc)
    ts, us = solve_fixed(orc.rhs, orc.u0, (0.0, 0.1), dt, alg, saveat=sv)
    assert sol.retcode == "Success" and len(us) == len(sv) and sol.stats["naccept"] == 15
    for k in range(len(sv)):
        np.testing.assert_allclose(sol.u[k], us[k], rtol=0, atol=1e-11)
    with pytest.raises(capi.MolError):
        mol_b200.solve(prob, A, dt=dt, adaptive=False, saveat=[0.05, 0.2])          # outside [t0, t1]


def test_step_to_ping_pong_equals_in_place_steps_and_reinit_drops_fsal():
    """mol_rk_step_to (two arrays, no state copy) against mol_rk_step (in place); a caller that rewrites u between steps
    gets a fresh k1 (continuation check / mol_rk_reinit) instead of the stale FSAL stage."""
    import torch
    prob = mol_b200.discretize(*examples.brusselator_2d(64))
    dev = torch.device("cuda", 0)
    n = prob.plan.state_len
    u0 = torch.from_numpy(prob.u0).to(dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    dt = 1e-6
    a, b = u0.clone(), torch.empty_like(u0)
    rk = capi.RK(prob.plan, "tsit5", 1e-6, 1e-3)
    t = 0.0
    for k in range(4):                                   # a -> b -> a -> b -> a
        src, dst = (a, b) if k % 2 == 0 else (b, a)
        t, _, s = rk.step_to(src.data_ptr(), dst.data_ptr(), t, dt, adaptive=False, stream=st)
        assert s.accepted == 1 and s.nf == (7 if k == 0 else 6)          # FSAL reused on continuation
    ref = u0.clone()
    rk2 = capi.RK(prob.plan, "tsit5", 1e-6, 1e-3)
    t2 = 0.0
    for k in range(4):
        t2, _, _ = rk2.step(ref.data_ptr(), t2, dt, adaptive=False, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(a, ref) and t == t2
    # the caller rewrites u in place: same pointer, same time -> must say so
    ref.mul_(1.5)
    rk2.reinit()
    t3, _, s = rk2.step(ref.data_ptr(), t2, dt, adaptive=False, stream=st)
    assert s.nf == 7
    fresh = capi.RK(prob.plan, "tsit5", 1e-6, 1e-3)
    chk = (a * 1.5).clone()
    fresh.step(chk.data_ptr(), t2, dt, adaptive=False, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(chk, ref)
    for r in (rk, rk2, fresh):
        r.close()


def test_persistent_solver_kernel_equals_host_driven_loop(monkeypatch):
    """Small problems are integrated by one launch of the persistent single-CTA kernel; MOL_RK_PERSISTENT=0 selects the
    host-driven loop.  Same step sequence, same saved states (config 1 and a fixed-step WENO solve)."""
    for mk, kw in ((lambda: examples.heat_1d_dirichlet(dx=0.01), dict(saveat=[0.0, 0.0137, 0.2, 0.731, 1.0], abstol=1e-8, reltol=1e-8)),
                   (lambda: examples.advection_1d_periodic(dx=0.02, scheme=mol_b200.WENOScheme(), tmax=0.1),
                    dict(alg=mol_b200.SSPRK33(), dt=0.0071, adaptive=False, saveat=[0.0, 0.01, 0.05, 0.0999, 0.1]))):
        kw = dict(kw)
        alg = kw.pop("alg", mol_b200.Tsit5())
        sols = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("MOL_RK_PERSISTENT", mode)
            prob = mol_b200.discretize(*mk())
            l0 = prob.plan.launch_count()
            sols[mode] = mol_b200.solve(prob, alg, **kw)
            launches = prob.plan.launch_count() - l0
            assert sols[mode].retcode == "Success"
            assert (launches <= 1 + len(kw["saveat"])) == (mode == "1")           # one solver launch (+ unpack is not counted here)
        assert sols["1"].stats["naccept"] == sols["0"].stats["naccept"] and sols["1"].stats["nreject"] == sols["0"].stats["nreject"]
        for ua, ub in zip(sols["1"].u, sols["0"].u):
            np.testing.assert_allclose(ua, ub, rtol=0, atol=1e-10)


def test_queued_adaptive_solve_equals_host_driven_loop(monkeypatch):
    """Mid-size adaptive Tsit5 solves run with the step controller on the device (attempts queued as a captured graph,
    csrc/mol_rk.cu solve_queued); MOL_RK_QUEUED=0 selects the host-driven loop, which runs the same controller kernel
    after every attempt.  Same step sequence (accepted and rejected attempts), same states at the save points (dense
    output inside steps, served while the queued solve is on hold): tiled 2-D, table-driven 1-D, z-marching 3-D.
    Explicit steps at the stability limit amplify last-bit differences to the level of the solver tolerance within ~200
    steps (the oracle's integrator moves by 1e-3 under a 1e-15 perturbation of u0 on the 64^2 Brusselator), so the spans
    compared against fixed step counts are short; the long run at the end relies on the error norm being summed in a
    fixed order."""
    import torch
    DEV = torch.device("cuda", 0)
    cases = ((lambda: examples.brusselator_2d(64, tmax=8e-5), dict(dt=4e-5, saveat=[0.0, 1.3e-5, 4.4e-5, 8e-5]), (6, 2)),
             (lambda: examples.heat_1d_dirichlet(dx=1.0 / 1100, tmax=2e-5), dict(dt=1e-5, saveat=[0.0, 1.1e-5, 2e-5], abstol=1e-8, reltol=1e-8), (15, 2)),
             (lambda: examples.diffusion_reaction_3d(n=16, periodic=True, tmax=0.01), dict(dt=5e-3, saveat=[0.0, 7e-3, 0.01]), (3, 0)))
    for mk, kw, (nacc, nrej) in cases:
        sols = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("MOL_RK_QUEUED", mode)
            prob = mol_b200.discretize(*mk())
            assert prob.plan.state_len > 1024
            sols[mode] = mol_b200.solve(prob, mol_b200.Tsit5(), **kw)
            assert sols[mode].retcode == "Success"
            if mode == "1":          # one integrator, three solves: same array twice (captured attempt reused), then another array
                rk = capi.RK(prob.plan, "tsit5", kw.get("abstol", 1e-6), kw.get("reltol", 1e-3))
                t0, t1 = prob.tspan
                st = torch.cuda.current_stream(DEV).cuda_stream
                bufs = [torch.empty(prob.plan.state_len, dtype=torch.float64, device=DEV) for _ in range(2)]
                for b in (bufs[0], bufs[0], bufs[1]):
                    b.copy_(torch.from_numpy(prob.u0))
                    stats = rk.solve(b.data_ptr(), t0, t1, kw["dt"], True, stream=st)
                    assert stats.retcode == 0 and stats.naccept == sols[mode].stats["naccept"]
                    np.testing.assert_allclose(b.cpu().numpy(), sols[mode].u[-1], rtol=0, atol=1e-7)
                rk.close()
        a, b = sols["1"].stats, sols["0"].stats
        assert (a["naccept"], a["nreject"]) == (b["naccept"], b["nreject"]) == (nacc, nrej), (a, b)      # (counts: the oracle's)
        assert len(sols["1"].u) == len(sols["0"].u) == len(kw["saveat"])
        for ua, ub in zip(sols["1"].u, sols["0"].u):
            np.testing.assert_allclose(ua, ub, rtol=0, atol=1e-7 * max(1.0, float(np.max(np.abs(ub)))))
    # a long run (stability-limited steps, ~185 attempts, holds at interior save points).  The error norm is summed without
    # atomics (static tile assignment in the FIN sweep, one slot per CTA, fixed-order reduction), so a solve reproduces
    # itself bit for bit and the two loops -- same kernels, same controller -- land on the same states
    sols = {}
    for mode in ("1", "0", "1 again"):
        monkeypatch.setenv("MOL_RK_QUEUED", mode[0])
        prob = mol_b200.discretize(*examples.brusselator_2d(64, tmax=2e-3))
        sols[mode] = mol_b200.solve(prob, mol_b200.Tsit5(), saveat=[0.0, 3.3e-4, 1.0e-3, 1.9e-3, 2e-3])
        assert sols[mode].retcode == "Success" and 150 <= sols[mode].stats["naccept"] <= 220
    for ua, ub, uc in zip(sols["1"].u, sols["0"].u, sols["1 again"].u):
        assert np.array_equal(ua, uc)                                    # run-to-run reproducibility
        np.testing.assert_allclose(ua, ub, rtol=0, atol=1e-9)
    assert sols["1"].stats == sols["0"].stats == sols["1 again"].stats


@pytest.mark.parametrize("alg", ["ssprk33", "tsit5"])
def test_fused_stage_loader_2d_many_tiles(alg):
    """Stage-combine-on-load across many tiles (128-bit loader on interior tiles, scalar loader + periodic wrap on
    edge tiles): 5 fixed steps of the Brusselator at N = 200 against the oracle's integrator."""
    from oracle.rk import solve_fixed
    sys_, disc = examples.brusselator_2d(200, tmax=1e-6)
    prob = mol_b200.discretize(sys_, disc)
    A = {"ssprk33": mol_b200.SSPRK33(), "tsit5": mol_b200.Tsit5()}[alg]
    dt = 2e-7
    sol = mol_b200.solve(prob, A, dt=dt, adaptive=False)
    orc = oracle_for(sys_, disc)
    ts, us = solve_fixed(orc.rhs, orc.u0, (0.0, 1e-6), dt, alg)
    assert sol.retcode == "Success"
    np.testing.assert_allclose(sol.u[-1], us[-1], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("alg", ["ssprk33", "tsit5"])
def test_fused_stage_loader_3d_zmarch(alg):
    """Fused RK stages through the z-marching 3-D kernel (cooperative plane loader, PRE/FIN epilogues): 3 fixed steps
    of the 3-D diffusion-reaction problem against the oracle's integrator."""
    from oracle.rk import solve_fixed
    dt = 1e-6
    sys_, disc = examples.diffusion_reaction_3d(n=72, periodic=True, nz=40, tmax=3 * dt)
    prob = mol_b200.discretize(sys_, disc)
    A = {"ssprk33": mol_b200.SSPRK33(), "tsit5": mol_b200.Tsit5()}[alg]
    sol = mol_b200.solve(prob, A, dt=dt, adaptive=False)
    orc = oracle_for(sys_, disc)
    ts, us = solve_fixed(orc.rhs, orc.u0, (0.0, 3 * dt), dt, alg)
    assert sol.retcode == "Success"
    np.testing.assert_allclose(sol.u[-1], us[-1], rtol=1e-12, atol=1e-12)


def test_adaptive_tsit5_error_estimate_matches_oracle():
    """The PRE/FIN split of the embedded error estimate: same accepted/rejected step sequence as the oracle's Tsit5."""
    from oracle.rk import solve_tsit5
    sys_, disc = examples.brusselator_2d(64, tmax=0.05)
    prob = mol_b200.discretize(sys_, disc)
    sol = mol_b200.solve(prob, mol_b200.Tsit5(), abstol=1e-8, reltol=1e-8)
    orc = oracle_for(sys_, disc)
    ts, us, stats = solve_tsit5(orc.rhs, orc.u0, (0.0, 0.05), abstol=1e-8, reltol=1e-8)
    assert sol.stats["naccept"] == stats["naccept"] and sol.stats["nreject"] == stats["nreject"], (sol.stats, stats)
    np.testing.assert_allclose(sol.u[-1], us[-1], rtol=1e-7, atol=1e-8)     # 10 x the integration tolerance


@pytest.mark.parametrize("name", ["heat_dirichlet", "heat_neumann", "heat_robin", "burgers_weno", "burgers2d", "burgers2d_nu",
                                  "advection2d_weno", "fisher3d_dirichlet_z", "fisher3d_periodic"])
def test_solution_unpacking_matches_oracle_full_state(name):
    """mol_unpack (sol[u(t,x)] of the reference, interface/solution/timedep.jl:30-72): unknowns + boundary nodes rebuilt
    from the boundary conditions + zero corner nodes, on the device, against the oracle's full_state."""
    import torch
    sys_, disc = CASES[name]()
    prob = mol_b200.discretize(sys_, disc)
    orc = oracle_for(sys_, disc)
    rng = np.random.default_rng(3)
    dev = torch.device("cuda", 0)
    times = [0.0, 0.37]
    states = np.stack([orc.u0 + 0.05 * rng.standard_normal(orc.nstate) for _ in times])
    shape = prob.plan.grid_shape(len(prob.program.axes))
    nodes = int(np.prod(shape))
    full = torch.empty((len(times), prob.plan.nvar, nodes), dtype=torch.float64, device=dev)
    prob.plan.unpack(full.data_ptr(), torch.from_numpy(states).to(dev).data_ptr(), times)
    torch.cuda.synchronize()
    got = full.cpu().numpy()
    for k, t in enumerate(times):
        ref = orc.full_state(states[k], t)
        for v in range(prob.plan.nvar):
            want = np.asarray(ref[v]).reshape(-1, order="F")
            scale = max(1.0, float(np.max(np.abs(want))))
            assert np.max(np.abs(got[k, v] - want)) <= 1e-12 * scale, (name, k, v)


def test_solution_indexing_full_grid():
    """sol[u(t,x)] has the time axis first and covers the whole grid; sol[x] / sol[t] return the grids
    (docs/src/tutorials/heat.md:43-58)."""
    sys_, disc = examples.heat_1d_dirichlet(dx=0.1)
    prob = mol_b200.discretize(sys_, disc)
    sol = mol_b200.solve(prob, mol_b200.Tsit5(), saveat=0.2)
    x = sol[prob.program.axes[0].sym]
    U = sol[sys_.dvs[0]]
    assert U.shape == (len(sol.t), len(x)) and len(x) == 11
    np.testing.assert_allclose(U[:, 0], np.exp(-sol.t), rtol=0, atol=1e-13)                  # Dirichlet data at x = 0
    np.testing.assert_allclose(U[:, -1], np.exp(-sol.t) * np.cos(1.0), rtol=0, atol=1e-13)   # and at x = 1
    assert np.max(np.abs(U - np.exp(-sol.t)[:, None] * np.cos(x)[None, :])) <= 0.01
    np.testing.assert_array_equal(sol.interior(sys_.dvs[0]), U[:, 1:-1])
