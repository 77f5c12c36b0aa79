"""Problem definitions mirrored from the reference's tests / docs / benchmark suite, written with the host-side mirror
API.  Test fixtures (they live under tests/, not in the product package): used by the tests, smoke() and bench.py, always
with synthetic inputs.  Import as `import problems` after `import _mol_import` (which puts tests/ on sys.path)."""
import numpy as np
import sympy as sp

from mol_b200.interface import (Differential, Eq, Interval, MOLFiniteDifference, PDESystem, UpwindScheme,
                        WENOScheme, ifelse)


def brusselator_2d(N, approx_order=2, doc_variant=False, tmax=11.5):
    """test/Brusselator/brusselator_eq.jl:8-63 (dx = dy = 1/N, periodic, alpha = 10).
    doc_variant: the system of docs/src/generated/bruss_code.md (same equations; the forcing term
    vanishes on the 4x4 doc grid)."""
    x, y, t = sp.symbols("x y t")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dxx, Dyy = Differential(t), Differential(x) ** 2, Differential(y) ** 2
    U, V = u(x, y, t), v(x, y, t)
    alpha = 10.0
    f = ifelse(sp.And((x - 0.3) ** 2 + (y - 0.6) ** 2 <= 0.1 ** 2, t >= 1.1), 5.0, 0.0)
    eqs = [Eq(Dt(U), 1.0 + V * U ** 2 - 4.4 * U + alpha * (Dxx(U) + Dyy(U)) + f),
           Eq(Dt(V), 3.4 * U - V * U ** 2 + alpha * (Dxx(V) + Dyy(V)))]
    bcs = [Eq(u(x, y, 0), 22 * (y * (1 - y)) ** 1.5), Eq(u(0, y, t), u(1, y, t)), Eq(u(x, 0, t), u(x, 1, t)),
           Eq(v(x, y, 0), 27 * (x * (1 - x)) ** 1.5), Eq(v(0, y, t), v(1, y, t)), Eq(v(x, 0, t), v(x, 1, t))]
    dom = [Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0), Interval(t, 0.0, tmax)]
    sys_ = PDESystem(eqs, bcs, dom, [x, y, t], [U, V], name="brusselator")
    disc = MOLFiniteDifference({x: 1.0 / N, y: 1.0 / N}, t, approx_order=approx_order)
    return sys_, disc


def heat_1d_dirichlet(dx=0.01, approx_order=2, tmax=1.0):
    """docs/src/tutorials/heat.md:19-41: u_t = u_xx on [0,1], u(t,0)=e^-t, u(t,1)=e^-t cos 1, u(0,x)=cos x."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dxx = Differential(t), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.cos(x)), Eq(u(t, 0), sp.exp(-t)), Eq(u(t, 1), sp.exp(-t) * sp.cos(1))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=approx_order)


def heat_1d_neumann(dx=0.1, approx_order=2, tmax=1.0):
    """docs/src/tutorials/heat.md:60-97 / test/Diffusion/MOL_1D_Linear_Diffusion.jl:179-251:
    Neumann at both ends, exact solution e^-t cos x on [0, 1]."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx, Dxx = Differential(t), Differential(x), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.cos(x)), Eq(Dx(u(t, 0)), 0.0), Eq(Dx(u(t, 1)), -sp.exp(-t) * sp.sin(1))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_neumann")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=approx_order)


def heat_1d_robin(dx=0.05, approx_order=2, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:374-428: Robin BCs on [-1, 1], exact e^-t sin x."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx, Dxx = Differential(t), Differential(x), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.sin(x)),
           Eq(u(t, -1.0) + 3 * Dx(u(t, -1.0)), sp.exp(-t) * (sp.sin(-1.0) + 3 * sp.cos(-1.0))),
           Eq(u(t, 1.0) + Dx(u(t, 1.0)), sp.exp(-t) * (sp.sin(1.0) + sp.cos(1.0)))]
    dom = [Interval(t, 0.0, tmax), Interval(x, -1.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_robin")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=approx_order)


def burgers_1d(dx=0.05, scheme=None, grid=None, tmax=1.0):
    """test/Burgers/burgers_eq.jl:6-54: u_t = -u u_x on [0,1], exact x/(t+1)."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    eq = Eq(Dt(u(t, x)), -u(t, x) * Dx(u(t, x)))
    bcs = [Eq(u(0, x), x), Eq(u(t, 0.0), 0.0), Eq(u(t, 1.0), 1.0 / (t + 1.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="burgers")
    spec = dx if grid is None else grid
    return sys_, MOLFiniteDifference({x: spec}, t, advection_scheme=scheme or UpwindScheme())


def advection_1d_periodic(dx=0.02, scheme=None, tmax=1.0, L=2.0):
    """benchmark/weno/problems.jl:7-30: u_t = -u_x on [0,2] periodic, IC sinpi(x)."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    eq = Eq(Dt(u(t, x)), -Dx(u(t, x)))
    bcs = [Eq(u(0, x), sp.sin(sp.pi * x)), Eq(u(t, 0.0), u(t, L))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, L)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="advection")
    return sys_, MOLFiniteDifference({x: dx}, t, advection_scheme=scheme or UpwindScheme())


def nonlinear_diffusion_1d(dx=0.05, approx_order=2, tmax=1.0):
    """test/Nonlinear_Diffusion/MOL_1D_NonLinear_Diffusion.jl: u_t = Dx(u^-2 ... ) family; here the
    c = 50 travelling-wave case  u_t = Dx(u^2 * Dx(u))-like coefficient a(u) = 1/(u^2) is avoided;
    uses a(u) = u^2 with Dirichlet data from the manufactured solution u = sqrt(x + 2 t + 1)."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    exact = sp.sqrt(x + 2 * t + 1)       # u^2 u_x = 1/2 * sqrt(..) ; Dx(u^2 u_x) = 1/(4 sqrt) ; u_t = 1/sqrt -> not exact
    eq = Eq(Dt(u(t, x)), Dx(u(t, x) ** 2 * Dx(u(t, x))))
    bcs = [Eq(u(0, x), sp.sqrt(x + 1)), Eq(u(t, 0.0), sp.sqrt(2 * t + 1)), Eq(u(t, 1.0), sp.sqrt(2 * t + 2))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="nonlinear_diffusion")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=approx_order)


def spherical_diffusion_1d(dr=0.05, tmax=0.5):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:479-537: u_t = 1/r^2 Dr(r^2 Dr u) on [0,1],
    exact sin(pi r)/(pi r) * exp(-pi^2 t)... here with Dr(u)(0)=0 and Dirichlet at r=1."""
    t, r = sp.symbols("t r")
    u = sp.Function("u")
    Dt, Dr = Differential(t), Differential(r)
    eq = Eq(Dt(u(t, r)), 1 / r ** 2 * Dr(r ** 2 * Dr(u(t, r))))
    bcs = [Eq(u(0, r), sp.sinc(sp.pi * r) if False else sp.Piecewise((1.0, r <= 0), (sp.sin(sp.pi * r) / (sp.pi * r), True))),
           Eq(Dr(u(t, 0.0)), 0.0), Eq(u(t, 1.0), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(r, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, r], [u(t, r)], name="spherical")
    return sys_, MOLFiniteDifference({r: dr}, t)


def burgers_2d(nx=32, ny=32, nu=1.0 / 80, grid_x=None, grid_y=None, tmax=0.5):
    """Config 3: 2-D viscous Burgers with UpwindScheme, x: Neumann/Neumann,
    y: Robin/Dirichlet, optional non-uniform grids."""
    t, x, y = sp.symbols("t x y")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    Dxx, Dyy = Differential(x) ** 2, Differential(y) ** 2
    U, V = u(t, x, y), v(t, x, y)
    eqs = [Eq(Dt(U), -U * Dx(U) - V * Dy(U) + nu * (Dxx(U) + Dyy(U))),
           Eq(Dt(V), -U * Dx(V) - V * Dy(V) + nu * (Dxx(V) + Dyy(V)))]
    bcs = [Eq(u(0, x, y), sp.sin(sp.pi * x) * sp.cos(sp.pi * y) * 0.5 + 0.2),
           Eq(v(0, x, y), sp.cos(sp.pi * x) * sp.sin(sp.pi * y) * 0.5 - 0.1),
           Eq(Dx(u(t, 0.0, y)), 0.0), Eq(Dx(u(t, 1.0, y)), 0.0),
           Eq(u(t, x, 0.0) + 0.5 * Dy(u(t, x, 0.0)), 0.2), Eq(u(t, x, 1.0), 0.2 - 0.5 * sp.sin(sp.pi * x) * sp.exp(-t)),
           Eq(Dx(v(t, 0.0, y)), 0.0), Eq(Dx(v(t, 1.0, y)), 0.0),
           Eq(v(t, x, 0.0) + 0.5 * Dy(v(t, x, 0.0)), -0.1), Eq(v(t, x, 1.0), -0.1)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x, y], [U, V], name="burgers2d")
    dxs = {x: (nx if grid_x is None else grid_x), y: (ny if grid_y is None else grid_y)}
    return sys_, MOLFiniteDifference(dxs, t, advection_scheme=UpwindScheme())


def diffusion_reaction_3d(n=16, periodic=True, tmax=0.1, D=1.0, nz=None):
    """Config 5: u_t = D lap(u) + u(1-u) on [0,1]^3, order-2 7-point stencil.
    nz: number of cells along z if different from n (same spacing h = 1/n; domain [0, nz/n] in z)."""
    t, x, y, z = sp.symbols("t x y z")
    u = sp.Function("u")
    U = u(t, x, y, z)
    Dt = Differential(t)
    lap = (Differential(x) ** 2)(U) + (Differential(y) ** 2)(U) + (Differential(z) ** 2)(U)
    eq = Eq(Dt(U), D * lap + U * (1 - U))
    zmax = 1.0 if nz is None else float(nz) / n
    ic = 0.5 + 0.25 * sp.sin(2 * sp.pi * x) * sp.cos(2 * sp.pi * y) * sp.sin(2 * sp.pi * z)
    bcs = [Eq(u(0, x, y, z), ic)]
    if periodic:
        bcs += [Eq(u(t, 0.0, y, z), u(t, 1.0, y, z)), Eq(u(t, x, 0.0, z), u(t, x, 1.0, z)),
                Eq(u(t, x, y, 0.0), u(t, x, y, zmax))]
    else:
        bcs += [Eq(u(t, 0.0, y, z), u(t, 1.0, y, z)), Eq(u(t, x, 0.0, z), u(t, x, 1.0, z)),
                Eq(u(t, x, y, 0.0), 0.5), Eq(u(t, x, y, zmax), 0.5)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0), Interval(z, 0.0, zmax)]
    sys_ = PDESystem([eq], bcs, dom, [t, x, y, z], [U], name="fisher3d")
    h = 1.0 / n
    return sys_, MOLFiniteDifference({x: h, y: h, z: h}, t)


def stretched_grid(a, b, n, amp=0.05):
    """benchmark/weno/grids.jl:13-19 (same generator as the reference's accuracy tests)."""
    xi = a + (b - a) * np.arange(n) / (n - 1)
    return xi + amp * np.sin(np.pi * (2 * (xi - a) / (b - a)))


def weno_burgers_periodic(dx=0.02, tmax=1.5):
    """benchmark/weno/problems.jl:32-50: u_t = -u u_x, periodic on [0,2], IC 1 + 0.25 sinpi(x), WENOScheme."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    eq = Eq(Dt(u(t, x)), -u(t, x) * Dx(u(t, x)))
    bcs = [Eq(u(0, x), 1.0 + 0.25 * sp.sin(sp.pi * x)), Eq(u(t, 0.0), u(t, 2.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="weno_burgers")
    return sys_, MOLFiniteDifference({x: dx}, t, advection_scheme=WENOScheme())


def advection_2d_periodic(n=32, scheme=None, tmax=0.5, ax=1.0, ay=0.5, approx_order=2, nu=0.0, grid_x=None, grid_y=None):
    """Config 4, 2-D form ("2-D WENO: tensor application per dim"): u_t = -ax u_x - ay u_y (+ nu lap u),
    periodic on [0,2]^2, IC sinpi(x) cospi(y)."""
    t, x, y = sp.symbols("t x y")
    u = sp.Function("u")
    U = u(t, x, y)
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    rhs = -ax * Dx(U) - ay * Dy(U)
    if nu:
        rhs = rhs + nu * ((Differential(x) ** 2)(U) + (Differential(y) ** 2)(U))
    bcs = [Eq(u(0, x, y), sp.sin(sp.pi * x) * sp.cos(sp.pi * y)),
           Eq(u(t, 0.0, y), u(t, 2.0, y)), Eq(u(t, x, 0.0), u(t, x, 2.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0), Interval(y, 0.0, 2.0)]
    sys_ = PDESystem([Eq(Dt(U), rhs)], bcs, dom, [t, x, y], [U], name="advection2d")
    h = 2.0 / n
    dxs = {x: (h if grid_x is None else grid_x), y: (h if grid_y is None else grid_y)}     # node vectors: non-uniform path
    return sys_, MOLFiniteDifference(dxs, t, advection_scheme=scheme or UpwindScheme(), approx_order=approx_order)


def heat_1d_neumann_pi(n=300, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:179-251 ("Test 03"): u_t = u_xx on [0, pi], homogeneous Neumann at both
    ends, u(0,x) = cos x, exact e^-t cos x; n nodes (the reference uses range(0, pi, length = 300))."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx, Dxx = Differential(t), Differential(x), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.cos(x)), Eq(Dx(u(t, 0)), 0.0), Eq(Dx(u(t, float(np.pi))), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, float(np.pi))]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_neumann_pi")
    return sys_, MOLFiniteDifference({x: int(n)}, t)


def heat_1d_robin_order4(dx=0.01, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:374-428 ("Test 05"): Robin BCs u + 3 u_x / 4 u + u_x on [-1, 1],
    approx_order = 4, exact e^-t sin x."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx, Dxx = Differential(t), Differential(x), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.sin(x)),
           Eq(u(t, -1.0) + 3 * Dx(u(t, -1.0)), sp.exp(-t) * (sp.sin(-1.0) + 3 * sp.cos(-1.0))),
           Eq(4 * u(t, 1.0) + Dx(u(t, 1.0)), sp.exp(-t) * (4 * sp.sin(1.0) + sp.cos(1.0)))]
    dom = [Interval(t, 0.0, tmax), Interval(x, -1.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_robin_o4")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=4)


def spherical_diffusion_order4(dr=0.1, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:479-537 ("Test 07"): u_t = 1/r^2 Dr(r^2 Dr u) on [0,1], Dr u(t,0) = 0,
    u(t,1) = e^-t sin 1, approx_order = 4, exact e^-t sin(r)/r."""
    t, r = sp.symbols("t r")
    u = sp.Function("u")
    Dt, Dr = Differential(t), Differential(r)
    eq = Eq(Dt(u(t, r)), 1 / r ** 2 * Dr(r ** 2 * Dr(u(t, r))))
    bcs = [Eq(u(0, r), sp.sin(r) / r), Eq(Dr(u(t, 0)), 0.0), Eq(u(t, 1), sp.exp(-t) * sp.sin(1))]
    dom = [Interval(t, 0.0, tmax), Interval(r, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, r], [u(t, r)], name="spherical_o4")
    return sys_, MOLFiniteDifference({r: dr}, t, approx_order=4)


def nonlinear_diffusion_travelling(dx=0.01, tmax=2.0, c=50.0, h=0.5):
    """test/Nonlinear_Diffusion/MOL_1D_NonLinear_Diffusion.jl:127-187 ("Test 01a"): u_t = Dx(u^2 Dx u) on [0,2], Dirichlet
    data and initial condition from the exact solution 0.5 (x + h)/sqrt(c - t)."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    exact = lambda tt, xx: 0.5 * (xx + h) / sp.sqrt(c - tt)
    eq = Eq(Dt(u(t, x)), Dx(u(t, x) ** 2 * Dx(u(t, x))))
    bcs = [Eq(u(0.0, x), exact(0.0, x)), Eq(u(t, 0.0), exact(t, 0.0)), Eq(u(t, 2.0), exact(t, 2.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="nonlinear_diffusion_travelling")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=2)


def convection_gaussian_periodic(dx=2.0 / 80, tmax=2.0, form="00"):
    """test/Convection/MOL_1D_Linear_Convection.jl:9-58 ("Test 00"): u_t = -u_x, periodic on [0,2], Gaussian pulse;
    the reference integrates with Euler(), dt = 0.025 (CFL 1: first-order upwind then shifts the pulse exactly).
    form: the same transport written the way Tests 00a (:59-107, u_t = u_x), 00b (:109-157, Dt u - Dx u ~ 0),
    00c (:159-206, Dt u + Dx u ~ 0), 01 (:208-255, source 0.001) and 02 (:257-314, speed as a number v) write it -- the
    upwind direction must follow the sign in each arrangement."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    asf = (0.5 / (0.2 * sp.sqrt(2.0 * 3.1415))) * sp.exp(-(x - 1.0) ** 2 / (2.0 * 0.2 ** 2))
    U = u(t, x)
    if form != "00":
        eq = {"00a": Eq(Dt(U), Dx(U)), "00b": Eq(Dt(U) - Dx(U), 0), "00c": Eq(Dt(U) + Dx(U), 0),
              "01": Eq(Dt(U), -Dx(U) + 0.001), "02": Eq(Dt(U), -1.0 * Dx(U))}[form]
        bcs = [Eq(u(0, x), asf), Eq(u(t, 0), u(t, 2))]
        dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0)]
        return PDESystem([eq], bcs, dom, [t, x], [U], name="convection_" + form), MOLFiniteDifference({x: dx}, t)
    eq = Eq(Dt(u(t, x)), -Dx(u(t, x)))
    bcs = [Eq(u(0, x), asf), Eq(u(t, 0), u(t, 2))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="convection")
    return sys_, MOLFiniteDifference({x: dx}, t)


def jittered_grid(a, b, n, amp=1e-3, seed=0):
    """range(a, b, length = n) with the interior nodes moved by +-amp at random, as the reference's non-uniform tests
    do (test/Diffusion_NU/MOL_1D_Linear_Diffusion_NonUniform.jl:38-41; the reference draws the signs with StableRNG(0),
    here NumPy's default_rng(seed): same construction, different signs)."""
    g = np.linspace(a, b, n)
    g[1:-1] += np.random.default_rng(seed).choice([amp, -amp], size=n - 2)
    return g


def heat_1d_dirichlet_pi(grid, approx_order=2, tmax=1.0):
    """test/Diffusion_NU/MOL_1D_Linear_Diffusion_NonUniform.jl:10-73 ("Test 00"): u_t = u_xx on [0, pi], u(t,0) = e^-t,
    u(t,pi) = -e^-t, u(0,x) = cos x, exact e^-t cos x; `grid` = dx, node count or node vector."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dxx = Differential(t), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.cos(x)), Eq(u(t, 0), sp.exp(-t)), Eq(u(t, float(np.pi)), -sp.exp(-t))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, float(np.pi))]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_dirichlet_pi")
    return sys_, MOLFiniteDifference({x: grid}, t, approx_order=approx_order)


def heat_1d_dirichlet_neumann_pi(grid, tmax=1.0):
    """test/Diffusion_NU/MOL_1D_Linear_Diffusion_NonUniform.jl:290-348 ("Test 04"): u(t,0) = 0, Dx u(t,pi) = -e^-t,
    u(0,x) = sin x, exact e^-t sin x."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx, Dxx = Differential(t), Differential(x), Differential(x) ** 2
    eq = Eq(Dt(u(t, x)), Dxx(u(t, x)))
    bcs = [Eq(u(0, x), sp.sin(x)), Eq(u(t, 0), 0.0), Eq(Dx(u(t, float(np.pi))), -sp.exp(-t))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, float(np.pi))]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_dirichlet_neumann_pi")
    return sys_, MOLFiniteDifference({x: grid}, t)


def diffusion_2d_dirichlet(dx=0.1, dy=0.2, approx_order=4, tmax=2.0):
    """test/2D_Diffusion/MOL_2D_Diffusion.jl:8-72 ("Test 00"): u_t = u_xx + u_yy on [0,2]^2, Dirichlet data and initial
    condition from the exact solution e^(x+y) cos(x + y + 4t), approx_order = 4."""
    t, x, y = sp.symbols("t x y")
    u = sp.Function("u")
    U = u(t, x, y)
    exact = lambda tt, xx, yy: sp.exp(xx + yy) * sp.cos(xx + yy + 4 * tt)
    eq = Eq(Differential(t)(U), (Differential(x) ** 2)(U) + (Differential(y) ** 2)(U))
    bcs = [Eq(u(0.0, x, y), exact(0.0, x, y)), Eq(u(t, 0.0, y), exact(t, 0.0, y)), Eq(u(t, 2.0, y), exact(t, 2.0, y)),
           Eq(u(t, x, 0.0), exact(t, x, 0.0)), Eq(u(t, x, 2.0), exact(t, x, 2.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0), Interval(y, 0.0, 2.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x, y], [U], name="diffusion2d")
    return sys_, MOLFiniteDifference({x: dx, y: dy}, t, approx_order=approx_order)


def three_species_2d(nx=24, ny=20, k=2.0):
    """A + B <-> C reaction-diffusion with a rate parameter k, periodic in x, homogeneous Neumann in y: three coupled
    equations, one parameter (exercises nvar = 3, `p`, mixed boundary types; cf. the multi-species systems of
    test/Components/interiormap_test.jl:8-40)."""
    t, x, y = sp.symbols("t x y")
    kk = sp.Symbol("k")
    a, b, c = sp.Function("a"), sp.Function("b"), sp.Function("c")
    A, B, Cc = a(t, x, y), b(t, x, y), c(t, x, y)
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    lap = lambda U: (Dx ** 2)(U) + (Dy ** 2)(U)
    eqs = [Eq(Dt(A), 0.1 * lap(A) - kk * A * B), Eq(Dt(B), 0.2 * lap(B) - kk * A * B + 0.5 * Cc),
           Eq(Dt(Cc), 0.05 * lap(Cc) + kk * A * B - 0.5 * Cc)]
    bcs = []
    for f, ic in ((a, 1 + 0.1 * sp.sin(2 * sp.pi * x)), (b, 0.5 + 0.1 * sp.cos(2 * sp.pi * x) * y), (c, 0.1 * y * (1 - y))):
        bcs += [Eq(f(0, x, y), ic), Eq(f(t, 0.0, y), f(t, 1.0, y)), Eq(Dy(f(t, x, 0.0)), 0.0), Eq(Dy(f(t, x, 1.0)), 0.0)]
    dom = [Interval(t, 0.0, 1.0), Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x, y], [A, B, Cc], ps=[(kk, k)], name="three_species")
    return sys_, MOLFiniteDifference({x: 1.0 / nx, y: 1.0 / ny}, t)


def advection_diffusion_robin_param(dx=0.02, a=0.2, b=1.5, tmax=1.0):
    """u_t = a u_xx - b u_x with parameters a, b that also appear in the boundary conditions: Dirichlet data b e^-t at
    x = 0 and the Robin condition u_x + a u = 0 at x = 1, whose coefficient is a parameter (ghost rule with an
    expression coefficient, `ghostx`)."""
    t, x = sp.symbols("t x")
    pa, pb = sp.symbols("a b")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    eq = Eq(Dt(u(t, x)), pa * (Dx ** 2)(u(t, x)) - pb * Dx(u(t, x)))
    bcs = [Eq(u(0, x), sp.exp(-10 * (x - 0.5) ** 2)), Eq(u(t, 0.0), pb * sp.exp(-t)), Eq(Dx(u(t, 1.0)) + pa * u(t, 1.0), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], ps=[(pa, a), (pb, b)], name="adv_diff_robin_param")
    return sys_, MOLFiniteDifference({x: dx}, t)


def heat_2d_robin_time_dependent(nx=24, ny=20, tmax=1.0):
    """2-D heat equation with a Robin condition whose coefficient varies in time and along the wall,
    u_x + (1 + 0.5 sin t + y) u = e^-t at x = 1; Neumann / Dirichlet elsewhere."""
    t, x, y = sp.symbols("t x y")
    u = sp.Function("u")
    U = u(t, x, y)
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    eq = Eq(Dt(U), (Dx ** 2)(U) + (Dy ** 2)(U))
    bcs = [Eq(u(0, x, y), sp.cos(sp.pi * x) * sp.sin(sp.pi * y) + 1),
           Eq(Dx(u(t, 0.0, y)), 0.0), Eq(Dx(u(t, 1.0, y)) + (1 + 0.5 * sp.sin(t) + y) * u(t, 1.0, y), sp.exp(-t)),
           Eq(u(t, x, 0.0), 1.0), Eq(u(t, x, 1.0), 1.0 + 0.2 * sp.sin(t) * x)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x, y], [U], name="heat2d_robin_t")
    return sys_, MOLFiniteDifference({x: 1.0 / nx, y: 1.0 / ny}, t)


# ---- variables on different domains joined by interface boundary conditions ----------------------------------------
def diffusion_two_domains(l=10, tmax=1.0, approx_order=2):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:887-930 (Test 14): c1 on [0, 0.5], c2 on [0.5, 1], joined by
    c1(t, 0.5) ~ c2(t, 0.5); homogeneous Dirichlet data at the outer ends; `l` points per domain."""
    t, x1, x2 = sp.symbols("t x1 x2")
    c1, c2 = sp.Function("c1"), sp.Function("c2")
    Dt, Dxx1, Dxx2 = Differential(t), Differential(x1) ** 2, Differential(x2) ** 2
    eqs = [Eq(Dt(c1(t, x1)), Dxx1(c1(t, x1))), Eq(Dt(c2(t, x2)), Dxx2(c2(t, x2)))]
    bcs = [Eq(c1(0, x1), -x1 * (x1 - 1) * sp.sin(x1)), Eq(c2(0, x2), x2 * (x2 - 1) * sp.sin(x2)),
           Eq(c1(t, 0), 0), Eq(c1(t, 0.5), c2(t, 0.5)), Eq(c2(t, 1), 0)]
    dom = [Interval(t, 0.0, tmax), Interval(x1, 0.0, 0.5), Interval(x2, 0.5, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x1, x2], [c1(t, x1), c2(t, x2)], name="diffusion_two_domains")
    return sys_, MOLFiniteDifference({x1: int(l), x2: int(l)}, t, approx_order=approx_order)


def right_cluster_grid(a, b, n, ratio=1000.0):
    """test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl:57-65: cells shrink geometrically towards b."""
    m = n - 1
    r = ratio ** (1.0 / (m - 1))
    dx = (r ** np.arange(m))[::-1].copy()
    dx *= (b - a) / dx.sum()
    x = a + np.concatenate([[0.0], np.cumsum(dx)])
    x[-1] = b
    return x


def one_sided_cluster_grid(a, b, n, ratio=1000.0):
    """same file :47-55: cells grow geometrically away from a."""
    m = n - 1
    r = ratio ** (1.0 / (m - 1))
    dx = r ** np.arange(m)
    dx *= (b - a) / dx.sum()
    x = a + np.concatenate([[0.0], np.cumsum(dx)])
    x[-1] = b
    return x


def advection_two_domains(x1grid=None, x2grid=None, v=1.0, v2=None, tmax=0.3, scheme=None, L=1.0):
    """solve_multi_domain_interface_advection (test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl:122-170; with
    WENOScheme: test/Convection_WENO/MOL_1D_WENO_NU_Interface.jl): u1_t = -v u1_x1 on x1grid, u2_t = -v2 u2_x2 on
    x2grid, u1(t, end) ~ u2(t, start); the inflow end carries the exact translating sine, the outflow end Dx u = 0.
    Grids: node vectors, or a float step / int point count per domain."""
    if x1grid is None:
        x1grid = right_cluster_grid(0.0, 0.5, 51, 500.0)
    if x2grid is None:
        x2grid = one_sided_cluster_grid(0.5, 1.0, 51, 500.0)
    v2 = v if v2 is None else v2
    t, x1, x2 = sp.symbols("t x1 x2")
    u1, u2 = sp.Function("u1"), sp.Function("u2")
    Dt, Dx1, Dx2 = Differential(t), Differential(x1), Differential(x2)

    def ends(g, a, b):
        return (float(g[0]), float(g[-1])) if np.ndim(g) > 0 else (a, b)
    a1, b1 = ends(x1grid, 0.0, 0.5)
    a2, b2 = ends(x2grid, 0.5, 1.0)
    exact = lambda xx, tt: sp.sin(2 * sp.pi * (xx - v * tt) / L)
    eqs = [Eq(Dt(u1(t, x1)), -v * Dx1(u1(t, x1))), Eq(Dt(u2(t, x2)), -v2 * Dx2(u2(t, x2)))]
    bcs = [Eq(u1(0, x1), exact(x1, 0)), Eq(u2(0, x2), exact(x2, 0)), Eq(u1(t, b1), u2(t, a2))]
    if v >= 0:
        bcs += [Eq(u1(t, a1), exact(a1, t)), Eq(Dx2(u2(t, b2)), 0.0)]
    else:
        bcs += [Eq(u2(t, b2), exact(b2, t)), Eq(Dx1(u1(t, a1)), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x1, a1, b1), Interval(x2, a2, b2)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x1, x2], [u1(t, x1), u2(t, x2)], name="advection_two_domains")
    return sys_, MOLFiniteDifference({x1: x1grid, x2: x2grid}, t, advection_scheme=scheme or UpwindScheme())


def advection_chained_domains(grids=None, v=1.0, tmax=0.1, scheme=None, L=1.0):
    """solve_chained_interface_advection (same file :172-215): four domains joined end to end, positive wind, exact
    inflow at the first lower end, Dx u = 0 at the last upper end."""
    if grids is None:
        edges = [0.0, 0.25, 0.5, 0.75, 1.0]
        grids = [right_cluster_grid(edges[0], edges[1], 21, 50.0), one_sided_cluster_grid(edges[1], edges[2], 26, 80.0),
                 np.linspace(edges[2], edges[3], 16), right_cluster_grid(edges[3], edges[4], 23, 30.0)]
    t = sp.Symbol("t")
    xs = sp.symbols("x1:%d" % (len(grids) + 1))
    us = [sp.Function("u%d" % (k + 1)) for k in range(len(grids))]
    Dt = Differential(t)
    exact = lambda xx, tt: sp.sin(2 * sp.pi * (xx - v * tt) / L)
    eqs = [Eq(Dt(us[k](t, xs[k])), -v * Differential(xs[k])(us[k](t, xs[k]))) for k in range(len(grids))]
    bcs = [Eq(us[k](0, xs[k]), exact(xs[k], 0)) for k in range(len(grids))]
    bcs += [Eq(us[k](t, float(grids[k][-1])), us[k + 1](t, float(grids[k + 1][0]))) for k in range(len(grids) - 1)]
    bcs += [Eq(us[0](t, float(grids[0][0])), exact(float(grids[0][0]), t)),
            Eq(Differential(xs[-1])(us[-1](t, float(grids[-1][-1]))), 0.0)]
    dom = [Interval(t, 0.0, tmax)] + [Interval(xs[k], float(grids[k][0]), float(grids[k][-1])) for k in range(len(grids))]
    sys_ = PDESystem(eqs, bcs, dom, [t] + list(xs), [us[k](t, xs[k]) for k in range(len(grids))], name="advection_chain")
    return sys_, MOLFiniteDifference({xs[k]: grids[k] for k in range(len(grids))}, t,
                                     advection_scheme=scheme or UpwindScheme())


def sinus_stretched_grid(a, b, n, amp=0.15):
    """test/Convection_WENO/MOL_1D_WENO_NU_Interface.jl:8-13: xi + amp sinpi(2 (xi - a) / (b - a)) on uniform xi."""
    xi = a + (b - a) * np.arange(n) / (n - 1)
    return xi + amp * np.sin(np.pi * (2 * (xi - a) / (b - a)))


def weno_pulse_two_domains(n1=41, n2=61, tmax=0.5, scheme=None):
    """test/Convection_WENO/MOL_1D_WENO_NU_Interface.jl:51-115: a Gaussian pulse advected across the interface of two
    deliberately mismatched non-uniform grids (u1 on [0, 1], u2 on [1, 2]); exact solution pulse(x, t)."""
    t, x1, x2 = sp.symbols("t x1 x2")
    u1, u2 = sp.Function("u1"), sp.Function("u2")
    Dt, Dx1, Dx2 = Differential(t), Differential(x1), Differential(x2)
    pulse = lambda xx, tt: sp.exp(-((xx - tt) - 0.7) ** 2 / (2 * 0.1 ** 2))
    eqs = [Eq(Dt(u1(t, x1)), -Dx1(u1(t, x1))), Eq(Dt(u2(t, x2)), -Dx2(u2(t, x2)))]
    bcs = [Eq(u1(0, x1), pulse(x1, 0.0)), Eq(u2(0, x2), pulse(x2, 0.0)), Eq(u1(t, 0.0), pulse(0.0, t)),
           Eq(u1(t, 1.0), u2(t, 1.0)), Eq(Dx2(u2(t, 2.0)), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x1, 0.0, 1.0), Interval(x2, 1.0, 2.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x1, x2], [u1(t, x1), u2(t, x2)], name="weno_pulse_two_domains")
    g1, g2 = sinus_stretched_grid(0.0, 1.0, n1, 0.03), sinus_stretched_grid(1.0, 2.0, n2, 0.04)
    return sys_, MOLFiniteDifference({x1: g1, x2: g2}, t, advection_scheme=scheme or WENOScheme())


def symmetric_cluster_grid(a, b, n, stretch=6.5):
    """test/Convection_NU/MOL_1D_Interface_Upwind_NonUniform.jl:38-45: nodes clustered at both ends."""
    xi = np.linspace(-1.0, 1.0, n)
    x = a + (b - a) * (np.sinh(stretch * xi) / np.sinh(stretch) + 1) / 2
    x[0], x[-1] = a, b
    return x


def chebyshev_nodes(a, b, n):
    """same file :30-36."""
    k = np.arange(1, n + 1)
    x = np.sort((a + b) / 2 + (b - a) / 2 * np.cos(np.pi * (2 * k - 1) / (2 * n)))
    x[0], x[-1] = a, b
    return x


def advection_periodic_speed(xgrid, v=1.0, tmax=0.4, ic=None, scheme=None):
    """solve_periodic_advection (same file :77-105): u_t = -v u_x, periodic, on a node vector; IC sin(2 pi x / L) or
    `ic(x)`."""
    xgrid = np.asarray(xgrid, dtype=float)
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    L = xgrid[-1] - xgrid[0]
    ic = (lambda xx: sp.sin(2 * sp.pi * xx / L)) if ic is None else ic
    eq = Eq(Differential(t)(u(t, x)), -v * Differential(x)(u(t, x)))
    bcs = [Eq(u(0, x), ic(x)), Eq(u(t, float(xgrid[0])), u(t, float(xgrid[-1])))]
    dom = [Interval(t, 0.0, tmax), Interval(x, float(xgrid[0]), float(xgrid[-1]))]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="advection_periodic_nu")
    return sys_, MOLFiniteDifference({x: xgrid}, t, advection_scheme=scheme or UpwindScheme())


def kdv_soliton(dx=0.4, tmax=1.0, alpha=6.0, beta=1.0):
    """test/Higher_Order/MOL_1D_HigherOrder.jl:94-152: u_t = -alpha u u_x - beta u_xxx on [-10, 10], three boundary
    conditions per end (u, Dx u, Dxx u from the single soliton 1/2 sech^2((x - t)/2))."""
    x, t = sp.symbols("x t")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    z = lambda xx, tt: (xx - tt) / 2
    sech2 = lambda q: sp.cosh(q) ** -2
    ua = lambda xx, tt: sech2(z(xx, tt)) / 2
    du = lambda xx, tt: sp.tanh(z(xx, tt)) * sech2(z(xx, tt)) / 2
    ddu = lambda xx, tt: (2 * sp.tanh(z(xx, tt)) ** 2 + sech2(z(xx, tt))) * sech2(z(xx, tt)) / 4
    eq = Eq(Dt(u(x, t)), -alpha * u(x, t) * Dx(u(x, t)) - beta * (Dx ** 3)(u(x, t)))
    bcs = [Eq(u(x, 0), ua(x, 0)), Eq(u(-10.0, t), ua(-10.0, t)), Eq(u(10.0, t), ua(10.0, t)),
           Eq(Dx(u(-10.0, t)), du(-10.0, t)), Eq(Dx(u(10.0, t)), du(10.0, t)),
           Eq((Dx ** 2)(u(-10.0, t)), ddu(-10.0, t)), Eq((Dx ** 2)(u(10.0, t)), ddu(10.0, t))]
    dom = [Interval(x, -10.0, 10.0), Interval(t, 0.0, tmax)]
    sys_ = PDESystem([eq], bcs, dom, [x, t], [u(x, t)], name="kdv")
    return sys_, MOLFiniteDifference({x: dx}, t)


def beam_with_velocity(dx=0.4, tmax=1.0, L=10.0, g=-9.81, EI=1.0, mu=1.0):
    """test/Higher_Order/MOL_1D_HigherOrder.jl:51-92 (Test 01): v ~ Dt(u), Dt(v) ~ -mu EI Dx^4 u + mu g, clamped at
    x = 0 (u = 0, v = 0), free at x = L (Dxx u = 0 and Dxxx u = 0: two conditions at one end), approx_order 4."""
    x, t = sp.symbols("x t")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx = Differential(t), Differential(x)
    eqs = [Eq(v(t, x), Dt(u(t, x))), Eq(Dt(v(t, x)), -mu * EI * (Dx ** 4)(u(t, x)) + mu * g)]
    bcs = [Eq(u(0, x), 0), Eq(v(0, x), 0), Eq(u(t, 0), 0), Eq(v(t, 0), 0),
           Eq((Dx ** 2)(u(t, L)), 0), Eq((Dx ** 3)(u(t, L)), 0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, L)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x], [u(t, x), v(t, x)], name="beam")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=4)


def anisotropic_diffusion_2d(nx=20, ny=16, periodic_y=False, tmax=0.1, kxy=0.5):
    """u_t = u_xx + u_yy + kxy u_xy: a mixed derivative (generate_mixed_rules, 2nd_order_mixed_deriv.jl:24-55; the reference
    only smoke-tests it, test/Mixed_Derivatives/MOL_Mixed_Deriv.jl).  Dirichlet data from exp(-t) sin(x + y) on the four
    walls, or periodic in y."""
    t, x, y = sp.symbols("t x y")
    u = sp.Function("u")
    U = u(t, x, y)
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    eq = Eq(Dt(U), (Dx ** 2)(U) + (Dy ** 2)(U) + kxy * Dx(Dy(U)))
    ex = lambda tt, xx, yy: sp.exp(-tt) * sp.sin(xx + yy) + 1
    Ly = 2 * sp.pi if periodic_y else 1.0
    bcs = [Eq(u(0, x, y), ex(0, x, y)), Eq(u(t, 0.0, y), ex(t, 0.0, y)), Eq(u(t, 1.0, y), ex(t, 1.0, y))]
    if periodic_y:
        bcs += [Eq(u(t, x, 0.0), u(t, x, float(Ly)))]
    else:
        bcs += [Eq(u(t, x, 0.0), ex(t, x, 0.0)), Eq(Dy(u(t, x, 1.0)), sp.exp(-t) * sp.cos(x + 1.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0), Interval(y, 0.0, float(Ly))]
    sys_ = PDESystem([eq], bcs, dom, [t, x, y], [U], name="anisotropic_diffusion")
    return sys_, MOLFiniteDifference({x: 1.0 / nx, y: float(Ly) / ny}, t)


def weno_mms_advection(xgrid, v=1.0, tmax=0.05, alpha=0.15):
    """test/Convection_WENO/MOL_1D_WENO_NU_Convergence.jl:60-76: u_t = -v u_x on a node vector spanning [0, 2 pi], Dirichlet
    data at both ends and initial condition from the manufactured solution sin(k) + alpha sin(2 k), k = 2 pi (x - v t) / L."""
    xgrid = np.asarray(xgrid, dtype=float)
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    L = 2 * sp.pi
    mms = lambda xx, tt: sp.sin(2 * sp.pi * (xx - v * tt) / L) + alpha * sp.sin(4 * sp.pi * (xx - v * tt) / L)
    x0, xL = float(xgrid[0]), float(xgrid[-1])
    eq = Eq(Differential(t)(u(t, x)), -v * Differential(x)(u(t, x)))
    bcs = [Eq(u(0.0, x), mms(x, 0.0)), Eq(u(t, x0), mms(x0, t)), Eq(u(t, xL), mms(xL, t))]
    dom = [Interval(t, 0.0, tmax), Interval(x, x0, xL)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="weno_mms")
    return sys_, MOLFiniteDifference({x: xgrid}, t, advection_scheme=WENOScheme())


def sinh_grid(a, b, n, beta=4.0):
    """same file :22-29: clustered at the centre."""
    xi = np.linspace(-1.0, 1.0, n)
    x = a + (b - a) * (np.sinh(beta * xi) / np.sinh(beta) + 1) / 2
    x[0], x[-1] = a, b
    return x


def tanh_grid(a, b, n, beta=2.0):
    """same file :31-38: clustered at the walls."""
    xi = np.linspace(-1.0, 1.0, n)
    x = a + (b - a) * (np.tanh(beta * xi) / np.tanh(beta) + 1) / 2
    x[0], x[-1] = a, b
    return x


def viscous_shock_grid(n=129):
    """same file :147-157: nodes equidistributed for the density 1 + 30 exp(-x^2 / (2 0.02^2)) on [-1, 1]."""
    xs = np.linspace(-1.0, 1.0, 5001)
    cdf = np.cumsum(1 + 30 * np.exp(-xs ** 2 / (2 * 0.02 ** 2)))
    cdf = (cdf - cdf[0]) / (cdf[-1] - cdf[0])
    out = []
    for l in np.linspace(0.0, 1.0, n):
        k = int(np.searchsorted(cdf, l, side="left"))
        if k <= 0:
            out.append(float(xs[0])); continue
        th = (l - cdf[k - 1]) / (cdf[k] - cdf[k - 1])
        out.append(float(xs[k - 1] + th * (xs[k] - xs[k - 1])))
    g = np.array(out)
    g[0], g[-1] = -1.0, 1.0
    return g


def viscous_shock(xgrid=None, nu=2.0e-3, tmax=1.0):
    """same file :137-186: u_t = -u u_x + nu u_xx (WENO advection + centred diffusion) holding the steady layer
    -tanh(x / (2 nu)) on a clustered grid, u(-1) = 1, u(1) = -1."""
    xgrid = viscous_shock_grid() if xgrid is None else np.asarray(xgrid, dtype=float)
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    eq = Eq(Dt(u(t, x)), -u(t, x) * Dx(u(t, x)) + nu * Dx(Dx(u(t, x))))
    bcs = [Eq(u(0.0, x), -sp.tanh(x / (2 * nu))), Eq(u(t, -1.0), 1.0), Eq(u(t, 1.0), -1.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, -1.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="viscous_shock")
    return sys_, MOLFiniteDifference({x: xgrid}, t, advection_scheme=WENOScheme())


def diffusion_two_independent_domains(l=100, tmax=1.0, approx_order=2):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:693-757 (Test 12): u(t, x) on [0, 1] and v(t, y) on [0, 2] in one system,
    not coupled; mixed Dirichlet / Neumann data from exp(-t) cos x and exp(-t) sin y."""
    t, x, y = sp.symbols("t x y")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    eqs = [Eq(Dt(u(t, x)), (Dx ** 2)(u(t, x))), Eq(Dt(v(t, y)), (Dy ** 2)(v(t, y)))]
    bcs = [Eq(u(0, x), sp.cos(x)), Eq(v(0, y), sp.sin(y)), Eq(u(t, 0), sp.exp(-t)), Eq(Dx(u(t, 1)), -sp.exp(-t) * sp.sin(1)),
           Eq(Dy(v(t, 0)), sp.exp(-t)), Eq(v(t, 2), sp.exp(-t) * sp.sin(2))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0), Interval(y, 0.0, 2.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x, y], [u(t, x), v(t, y)], name="two_independent_domains")
    return sys_, MOLFiniteDifference({x: 1.0 / (l - 1), y: 2.0 / (l - 1)}, t, approx_order=approx_order)


def heat_1d_robin_time_dependent(dx=0.01, approx_order=6, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:429-477 (Test 06): u_t = u_xx on [-1, 1] with the time-dependent Robin
    conditions t^2 u + 3 u_x and 4 u + t u_x (data from exp(-t) sin x), order 6."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    eq = Eq(Dt(u(t, x)), (Dx ** 2)(u(t, x)))
    bcs = [Eq(u(0, x), sp.sin(x)),
           Eq(t ** 2 * u(t, -1.0) + 3 * Dx(u(t, -1.0)), sp.exp(-t) * (t ** 2 * sp.sin(-1.0) + 3 * sp.cos(-1.0))),
           Eq(4 * u(t, 1.0) + t * Dx(u(t, 1.0)), sp.exp(-t) * (4 * sp.sin(1.0) + t * sp.cos(1.0)))]
    dom = [Interval(t, 0.0, tmax), Interval(x, -1.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="heat_robin_time_dependent")
    return sys_, MOLFiniteDifference({x: dx}, t, approx_order=approx_order)


def diffusion_two_variables_mixed_bcs(l=100, approx_order=2, tmax=1.0):
    """same file :599-657 (Test 10): u and v on one grid with opposite Dirichlet / Neumann ends (exp(-t) cos x, exp(-t) sin x)."""
    t, x = sp.symbols("t x")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx = Differential(t), Differential(x)
    eqs = [Eq(Dt(u(t, x)), (Dx ** 2)(u(t, x))), Eq(Dt(v(t, x)), (Dx ** 2)(v(t, x)))]
    bcs = [Eq(u(0, x), sp.cos(x)), Eq(v(0, x), sp.sin(x)), Eq(u(t, 0), sp.exp(-t)), Eq(Dx(u(t, 1)), -sp.exp(-t) * sp.sin(1)),
           Eq(Dx(v(t, 0)), sp.exp(-t)), Eq(v(t, 1), sp.exp(-t) * sp.sin(1))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x], [u(t, x), v(t, x)], name="two_variables_mixed_bcs")
    return sys_, MOLFiniteDifference({x: 1.0 / (l - 1)}, t, approx_order=approx_order)


def reaction_diffusion_parameters(dx=0.1, Dn=0.5, Dp=2.0, tmax=1.0):
    """same file :659-691 (Test 11): two species with parameter diffusivities Dn, Dp and the reaction +-u v."""
    t, x = sp.symbols("t x")
    pn, pp = sp.symbols("Dn Dp")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx = Differential(t), Differential(x)
    eqs = [Eq(Dt(u(t, x)), pn * (Dx ** 2)(u(t, x)) + u(t, x) * v(t, x)),
           Eq(Dt(v(t, x)), pp * (Dx ** 2)(v(t, x)) - u(t, x) * v(t, x))]
    bcs = [Eq(u(0, x), sp.sin(sp.pi * x / 2)), Eq(v(0, x), sp.sin(sp.pi * x / 2)),
           Eq(u(t, 0), 0.0), Eq(Dx(u(t, 1)), 0.0), Eq(v(t, 0), 0.0), Eq(Dx(v(t, 1)), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x], [u(t, x), v(t, x)], ps=[(pn, Dn), (pp, Dp)], name="reaction_diffusion_params")
    return sys_, MOLFiniteDifference({x: dx}, t)


def diffusion_variable_coefficient(N=11, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:131-177 (Test 02): u_t = Dx(D) Dx(u) + D Dxx(u) with the known coefficient
    D(t, x) = 0.999 + 0.001 t x (its derivative is expanded symbolically; the first-derivative term is upwinded on the sign
    of Dx(D) = 0.001 t); N points."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    D = 0.999 + 0.001 * t * x
    eq = Eq(Dt(u(t, x)), Dx(D) * Dx(u(t, x)) + D * (Dx ** 2)(u(t, x)))
    bcs = [Eq(u(0, x), -x * (x - 1) * sp.sin(x)), Eq(u(t, 0), 0.0), Eq(u(t, 1), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="diffusion_variable_coefficient")
    return sys_, MOLFiniteDifference({x: int(N)}, t)


def diffusion_with_ode(l=100, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:830-885 (Test 13): one diffusion equation with mixed BCs and one ODE
    Dt(v(t)) ~ -v(t) in the same system (exact: exp(-t) sin x and exp(-t))."""
    t, x = sp.symbols("t x")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx = Differential(t), Differential(x)
    eqs = [Eq(Dt(u(t, x)), (Dx ** 2)(u(t, x))), Eq(Dt(v(t)), -v(t))]
    bcs = [Eq(u(0, x), sp.sin(x)), Eq(v(0), 1), Eq(u(t, 0), 0), Eq(Dx(u(t, 1)), sp.exp(-t) * sp.cos(1))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x], [u(t, x), v(t)], name="diffusion_with_ode")
    return sys_, MOLFiniteDifference({x: 1.0 / (l - 1)}, t)


def nonlinear_diffusion_2d(dx=0.1, dy=0.2, tmax=2.0):
    """test/2D_Diffusion/MOL_2D_Diffusion.jl:75-150 (Test 01): u_t = Dx(a Dx u) + Dy(a Dy u) with
    a = sqrt(u^2 / exp(x + y)^2 + sin(x + y + 4 t)^2) -- the nonlinear Laplacian in both dimensions, coefficient depending
    on u, x, y and t -- Dirichlet data from exp(x + y) cos(x + y + 4 t) on [0, 2]^2."""
    t, x, y = sp.symbols("t x y")
    u = sp.Function("u")
    U = u(t, x, y)
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    ex = lambda tt, xx, yy: sp.exp(xx + yy) * sp.cos(xx + yy + 4 * tt)
    a = (U ** 2 / sp.exp(x + y) ** 2 + sp.sin(x + y + 4 * t) ** 2) ** 0.5
    eq = Eq(Dt(U), Dx(a * Dx(U)) + Dy(a * Dy(U)))
    bcs = [Eq(u(0.0, x, y), ex(0.0, x, y)), Eq(u(t, 0.0, y), ex(t, 0.0, y)), Eq(u(t, 2.0, y), ex(t, 2.0, y)),
           Eq(u(t, x, 0.0), ex(t, x, 0.0)), Eq(u(t, x, 2.0), ex(t, x, 2.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0), Interval(y, 0.0, 2.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x, y], [U], name="nonlinear_diffusion_2d")
    return sys_, MOLFiniteDifference({x: dx, y: dy}, t)


def advection_dirichlet_nu(xgrid, v=1.0, tmax=0.4, scheme=None):
    """solve_mms_advection (test/Convection_NU/MOL_1D_Linear_Convection_NonUniform.jl:77-110): u_t = -v u_x on a node
    vector with the exact translating sine sin(2 pi (x - v t) / L) as data at both ends."""
    xgrid = np.asarray(xgrid, dtype=float)
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    L = xgrid[-1] - xgrid[0]
    ex = lambda xx, tt: sp.sin(2 * sp.pi * (xx - v * tt) / L)
    x0, xL = float(xgrid[0]), float(xgrid[-1])
    eq = Eq(Differential(t)(u(t, x)), -v * Differential(x)(u(t, x)))
    bcs = [Eq(u(0.0, x), ex(x, 0.0)), Eq(u(t, x0), ex(x0, t)), Eq(u(t, xL), ex(xL, t))]
    dom = [Interval(t, 0.0, tmax), Interval(x, x0, xL)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="advection_dirichlet_nu")
    return sys_, MOLFiniteDifference({x: xgrid}, t, advection_scheme=scheme or UpwindScheme())


def advection_inflow_nu(xgrid, v=0.8, tmax=0.2, scheme=None, dx=None):
    """solve_inflow_advection (same file :116-153): inflow datum sin(2 pi t / L) at the upwind end, Dx u = 0 at the outflow end.
    dx: use the uniform step dx on [xgrid[0], xgrid[-1]] instead of the node vector."""
    xgrid = np.asarray(xgrid, dtype=float)
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    L = xgrid[-1] - xgrid[0]
    uL = lambda tt: sp.sin(2 * sp.pi * tt / L)
    exact0 = uL(-x / v) if v >= 0 else uL(-(L - x) / abs(v))
    inflow, outflow = (float(xgrid[0]), float(xgrid[-1])) if v >= 0 else (float(xgrid[-1]), float(xgrid[0]))
    eq = Eq(Differential(t)(u(t, x)), -v * Differential(x)(u(t, x)))
    bcs = [Eq(u(0.0, x), exact0), Eq(u(t, inflow), uL(t)), Eq(Differential(x)(u(t, outflow)), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, float(xgrid[0]), float(xgrid[-1]))]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="advection_inflow_nu")
    return sys_, MOLFiniteDifference({x: xgrid if dx is None else float(dx)}, t, advection_scheme=scheme or UpwindScheme())


def diffusion_driven_by_ode(l=30, tmax=1.0):
    """A field driven by a variable of t alone: Dt(u) ~ Dxx(u) + v(t) sin(pi x), Dt(v) ~ -v (one-way coupling; the reverse
    needs boundary values / integrals inside equations, which stay out of scope)."""
    t, x = sp.symbols("t x")
    u, v = sp.Function("u"), sp.Function("v")
    Dt, Dx = Differential(t), Differential(x)
    eqs = [Eq(Dt(u(t, x)), (Dx ** 2)(u(t, x)) + v(t) * sp.sin(sp.pi * x)), Eq(Dt(v(t)), -v(t))]
    bcs = [Eq(u(0, x), sp.sin(sp.pi * x)), Eq(v(0), 1), Eq(u(t, 0), 0), Eq(u(t, 1), 0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem(eqs, bcs, dom, [t, x], [u(t, x), v(t)], name="diffusion_driven_by_ode")
    return sys_, MOLFiniteDifference({x: 1.0 / (l - 1)}, t)


def heat_parameter_diffusivity(dx=1.0 / (5 * np.pi), D=10.0, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:87-129 (Test 01): u_t = D u_xx with the parameter D = 10; dx = 1 / (5 pi) does
    not divide [0, 1], so the reference appends the node x = 1 (discretize_vars.jl:224-229) and the grid becomes a node vector
    with a short last cell."""
    t, x, Dp = sp.symbols("t x D")
    u = sp.Function("u")
    eq = Eq(Differential(t)(u(t, x)), Dp * (Differential(x) ** 2)(u(t, x)))
    bcs = [Eq(u(0, x), -x * (x - 1) * sp.sin(x)), Eq(u(t, 0), 0.0), Eq(u(t, 1), 0.0)]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], ps=[(Dp, D)], name="heat_parameter_D")
    return sys_, MOLFiniteDifference({x: float(dx)}, t)


def spherical_diffusion_coefficient4(dr=0.1, tmax=1.0):
    """test/Diffusion/MOL_1D_Linear_Diffusion.jl:539-597 (Test 08): u_t = 4 / r^2 Dr(r^2 Dr u) on [0, 1] (a constant factor in
    front of the spherical Laplacian), Dr u(t, 0) = 0, u(t, 1) = exp(-4 t) sin 1; exact exp(-4 t) sin(r) / r."""
    t, r = sp.symbols("t r")
    u = sp.Function("u")
    Dt, Dr = Differential(t), Differential(r)
    eq = Eq(Dt(u(t, r)), 4 / r ** 2 * Dr(r ** 2 * Dr(u(t, r))))
    bcs = [Eq(u(0, r), sp.sin(r) / r), Eq(Dr(u(t, 0)), 0), Eq(u(t, 1), sp.exp(-4 * t) * sp.sin(1))]
    dom = [Interval(t, 0.0, tmax), Interval(r, 0.0, 1.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, r], [u(t, r)], name="spherical4")
    return sys_, MOLFiniteDifference({r: dr}, t)


def nonlinear_diffusion_inverse(dx=0.01, tmax=2.0, c=1.0, a=1.0):
    """test/Nonlinear_Diffusion/MOL_1D_NonLinear_Diffusion.jl:12-69 (Test 00): u_t = Dx(u^-1 Dx u) on [0, 2], Dirichlet data
    and initial condition from the exact solution 2 (c + t) / (a + x)^2."""
    t, x = sp.symbols("t x")
    u = sp.Function("u")
    Dt, Dx = Differential(t), Differential(x)
    exact = lambda tt, xx: 2.0 * (c + tt) / (a + xx) ** 2
    eq = Eq(Dt(u(t, x)), Dx(u(t, x) ** -1 * Dx(u(t, x))))
    bcs = [Eq(u(0.0, x), exact(0.0, x)), Eq(u(t, 0.0), exact(t, 0.0)), Eq(u(t, 2.0), exact(t, 2.0))]
    dom = [Interval(t, 0.0, tmax), Interval(x, 0.0, 2.0)]
    sys_ = PDESystem([eq], bcs, dom, [t, x], [u(t, x)], name="nonlinear_diffusion_inverse")
    return sys_, MOLFiniteDifference({x: dx}, t)
