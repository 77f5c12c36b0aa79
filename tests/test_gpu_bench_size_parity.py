"""Parity at the sizes that are benchmarked: the tiled kernels' non-uniform / cp.async / z-march
paths against looped C restatements (oracle/configs_ref.c, validated against the generic Python oracle at small sizes
in tests/test_cref_cpu.py) -- config 3 at 4097^2 (non-uniform, Neumann / Robin / Dirichlet), config 4 in 1-D at 2^20
nodes through the per-node properties the domain offers, config 5 at 1024^2 x 128 and one 1024^3 evaluation (8.6 GB
state: 64-bit indexing end to end)."""
import numpy as np
import pytest

import mol_b200
from mol_b200 import capi
import problems as examples
from oracle import cref

pytestmark = pytest.mark.gpu


def _gpu_rhs(prob, u, t):
    import torch
    dev = torch.device("cuda", prob.device)
    ud = torch.from_numpy(u).to(dev)
    dud = torch.empty_like(ud)
    prob.f(dud, ud, None, t)
    torch.cuda.synchronize()
    return dud.cpu().numpy()


@pytest.mark.parametrize("nonuniform", [True, False])
def test_config3_burgers_4097_matches_c_restatement(nonuniform):
    n = 4097
    if nonuniform:
        gx = 0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, n)) / np.tanh(2.0))
        gy = np.linspace(0, 1, n) ** 1.3
        sys_, disc = examples.burgers_2d(grid_x=gx, grid_y=gy)
    else:
        sys_, disc = examples.burgers_2d(nx=n - 1, ny=n - 1)
    prob = mol_b200.discretize(sys_, disc)
    assert prob.program.corebox is not None
    B = cref.Burgers2D(sys_, disc)
    assert B.nstate == prob.plan.state_len
    rng = np.random.default_rng(5)
    u = B.orc.u0 + 0.1 * rng.standard_normal(B.nstate)
    for t in (0.0, 0.37):
        ref = B.rhs(u, t, nthreads=cref.host_threads())
        got = _gpu_rhs(prob, u, t)
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), (nonuniform, t)


def test_config5_fisher3d_slab_1024x1024x128_matches_c_restatement():
    n, nz = 1024, 128
    sys_, disc = examples.diffusion_reaction_3d(n=n, periodic=True, nz=nz)
    prob = mol_b200.discretize(sys_, disc)
    u = np.random.default_rng(6).uniform(0.0, 1.0, prob.plan.state_len)
    P = n * n
    ref = cref.fisher3d_rhs_slab(u, u[-P:], u[:P], n, n, nz, 1.0 / n, nthreads=cref.host_threads())
    got = _gpu_rhs(prob, u, 0.0)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_config5_fisher3d_1024_cubed_on_one_gpu():
    """1024^3 = 2^30 unknowns, 8.6 GB per array: flat indices beyond 2^31 bytes / 2^30 elements through the tensor maps,
    the z-march ring and the stores.  Checked plane-wise against the C restatement (periodic images across the ends)."""
    import torch
    n = 1024
    free, _ = torch.cuda.mem_get_info()
    if free < 20 * 2 ** 30:
        pytest.skip("needs 20 GB of free device memory")
    sys_, disc = examples.diffusion_reaction_3d(n=n, periodic=True)
    prob = mol_b200.discretize(sys_, disc)
    assert prob.plan.state_len == n ** 3
    dev = torch.device("cuda", prob.device)
    g = torch.Generator(device=dev).manual_seed(3)
    ud = torch.rand(n ** 3, dtype=torch.float64, device=dev, generator=g)
    dud = torch.empty_like(ud)
    prob.f(dud, ud, None, 0.0)
    torch.cuda.synchronize()
    P = n * n
    U = ud.view(n, P)
    D = dud.view(n, P)
    for k0, k1 in ((0, 3), (510, 514), (1021, 1024)):          # slabs at both ends (periodic wrap) and across 2^31 bytes
        slab = U[k0:k1].cpu().numpy().reshape(-1)
        lo = U[(k0 - 1) % n].cpu().numpy()
        hi = U[k1 % n].cpu().numpy()
        ref = cref.fisher3d_rhs_slab(slab, lo, hi, n, n, k1 - k0, 1.0 / n)
        got = D[k0:k1].cpu().numpy().reshape(-1)
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), (k0, k1)


def test_config4_weno_1d_2pow20_properties():
    """2^20-node 1-D WENO5 (uniform and stretched, periodic): translation invariance by whole periods is exact for the
    uniform kernel, du of a constant state is 0, and the tiled kernel agrees with the table-driven kernel to 1e-12
    (the latter is checked against the oracle at small sizes)."""
    import torch
    n = 1 << 20
    for mk in (lambda: examples.advection_1d_periodic(dx=2.0 / n, scheme=mol_b200.WENOScheme()),
               lambda: examples.advection_1d_periodic(dx=examples.stretched_grid(0, 2, n + 1), scheme=mol_b200.WENOScheme())):
        prob = mol_b200.discretize(*mk())
        assert prob.program.corebox is not None
        u = prob.u0 + 0.05 * np.random.default_rng(8).standard_normal(prob.plan.state_len)
        prob.plan.set_option("kernel", capi.KERNEL_AUTO)
        tiled = _gpu_rhs(prob, u, 0.0)
        prob.plan.set_option("kernel", capi.KERNEL_GENERIC)
        generic = _gpu_rhs(prob, u, 0.0)
        prob.plan.set_option("kernel", capi.KERNEL_AUTO)
        assert np.max(np.abs(tiled - generic)) <= 1e-12 * np.max(np.abs(generic))
        const = _gpu_rhs(prob, np.full_like(u, 1.25), 0.0)
        assert np.max(np.abs(const)) <= 1e-9                     # exact 0 on a uniform grid, rounding of 1/h sums otherwise
