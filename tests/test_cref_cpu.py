"""The looped C restatements (oracle/bruss_ref.c, oracle/configs_ref.c) against the generic Python oracle
(oracle/discretize.py) at small sizes: they are the checkers the GPU suite and bench.py use at benchmark size."""
import numpy as np
import pytest

import _mol_import  # noqa: F401
import problems as examples
from oracle import cref
from oracle.discretize import OracleProblem


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("t", [0.0, 2.0])
def test_bruss_slab_restatement_matches_oracle(t):
    N, P = 16, 3                                          # three slabs stacked along y (weak-scaling layout)
    from mol_b200.distributed import stack_domain
    sys_, disc = stack_domain(*examples.brusselator_2d(N), P)
    orc = OracleProblem(sys_, disc)
    rng = np.random.default_rng(3)
    ug = rng.uniform(0.0, 3.0, orc.nstate)
    ref = orc.rhs(ug, t).reshape(2, N * P, N)
    U = ug.reshape(2, N * P, N)
    xg = np.arange(N + 1) / N
    for r in range(P):
        a, b = r * N, (r + 1) * N
        lo, hi = (a - 1) % (N * P), b % (N * P)
        yrows = (np.arange(a, b) + 1) / N
        got = cref.bruss_rhs_slab(U[:, a:b].reshape(-1), U[0, lo], U[0, hi], U[1, lo], U[1, hi], xg, yrows, N, N, t)
        assert _rel(got.reshape(2, N, N), ref[:, a:b]) <= 1e-13


def test_bruss_slab_equals_whole_domain_restatement():
    N = 24
    rng = np.random.default_rng(0)
    u = rng.uniform(0.0, 3.0, 2 * N * N)
    g = np.arange(N + 1) / N
    whole = cref.bruss_rhs(u, g, g, N, 2.0)
    U = u.reshape(2, N, N)
    slab = cref.bruss_rhs_slab(u, U[0, -1], U[0, 0], U[1, -1], U[1, 0], g, g[1:], N, N, 2.0)
    assert np.array_equal(whole, slab)


def test_fisher3d_slab_restatement_matches_oracle():
    n, nz = 8, 12
    sys_, disc = examples.diffusion_reaction_3d(n=n, periodic=True, nz=nz)
    orc = OracleProblem(sys_, disc)
    rng = np.random.default_rng(5)
    ug = rng.uniform(0.0, 1.0, orc.nstate)
    ref = orc.rhs(ug, 0.0).reshape(nz, n, n)
    U = ug.reshape(nz, n, n)
    for a, b in ((0, 5), (5, 12)):
        got = cref.fisher3d_rhs_slab(U[a:b].reshape(-1), U[(a - 1) % nz], U[b % nz], n, n, b - a, 1.0 / n)
        assert _rel(got.reshape(b - a, n, n), ref[a:b]) <= 1e-13


def test_host_threads_ignores_omp_num_threads(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    import os
    assert cref.host_threads() == len(os.sched_getaffinity(0))


@pytest.mark.parametrize("nonuniform", [False, True])
def test_burgers2d_restatement_matches_oracle(nonuniform):
    """Config 3 (upwind Burgers, Neumann / Robin / Dirichlet, optional non-uniform grids): the looped C evaluation with
    the oracle's row tables against the generic Python oracle."""
    if nonuniform:
        gx = 0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, 41)) / np.tanh(2.0))
        gy = np.linspace(0, 1, 37) ** 1.3
        sys_, disc = examples.burgers_2d(grid_x=gx, grid_y=gy)
    else:
        sys_, disc = examples.burgers_2d(nx=40, ny=36)
    B = cref.Burgers2D(sys_, disc)
    orc = B.orc
    assert B.nstate == orc.nstate
    rng = np.random.default_rng(2)
    for u in (orc.u0, orc.u0 + 0.3 * rng.standard_normal(orc.nstate)):
        for t in (0.0, 0.37):
            ref = orc.rhs(u, t)
            got = B.rhs(u, t)
            scale = float(np.max(orc.rhs_termscale(u, t)))
            assert np.max(np.abs(got - ref)) <= 1e-13 * scale
