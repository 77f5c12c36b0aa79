"""mol_plan_jac_sparsity (first half): the Jacobian pattern read off the stencil program must contain
every entry of the oracle's numerically differentiated Jacobian, and be exactly it where nothing is branch-dependent."""
import numpy as np
import pytest

import mol_b200
from mol_b200 import capi
from oracle.discretize import OracleProblem

from test_ir_semantics_cpu import CASES

EXACT = {"brusselator", "heat_neumann", "heat_robin", "heat_dirichlet_o4", "fisher3d_dirichlet_z", "edge_heat_neumann",
         "nu_heat_dirichlet", "diffusion2d_o4", "nonlinear_diffusion", "robin_time_dependent_2d", "mixed_derivative",
         "mixed_derivative_periodic_y"}


@pytest.mark.parametrize("name", ["brusselator", "heat_neumann", "heat_robin", "heat_dirichlet_o4", "burgers_upwind",
                                  "burgers_weno", "advection_weno_periodic", "advection_weno_stretched", "nonlinear_diffusion",
                                  "spherical", "burgers2d", "burgers2d_nu", "fisher3d_dirichlet_z", "edge_heat_neumann",
                                  "edge_burgers2d", "nu_heat_dirichlet", "diffusion2d_o4", "robin_parameter_coefficient",
                                  "robin_time_dependent_2d", "kdv_three_bcs_per_end", "beam_two_bcs_at_free_end",
                                  "mixed_derivative", "mixed_derivative_periodic_y"])
def test_jacobian_pattern_covers_numerical_jacobian(name):
    sys_, disc = CASES[name]()
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    colptr, rowval = plan.jac_sparsity()
    n = prog.nstate
    assert colptr[0] == 0 and colptr[-1] == len(rowval) and np.all(np.diff(colptr) >= 0)
    pattern = np.zeros((n, n), dtype=bool)
    for j in range(n):
        rows = rowval[colptr[j]:colptr[j + 1]]
        assert np.all(np.diff(rows) > 0)                      # sorted, no duplicates
        pattern[rows, j] = True
    orc = OracleProblem(sys_, disc)
    rng = np.random.default_rng(12)
    numeric = np.zeros((n, n), dtype=bool)
    for trial in range(2):                                    # two states: entries can vanish by accident at one
        u = np.abs(orc.u0 + 0.3 * rng.standard_normal(n)) + 0.2
        f0 = orc.rhs(u, 0.37)
        for j in range(n):
            h = 1e-6 * max(1.0, abs(u[j]))
            up = u.copy(); up[j] += h
            numeric[:, j] |= np.abs(orc.rhs(up, 0.37) - f0) > 1e-9 * h * max(1.0, float(np.max(np.abs(f0))))
    missing = numeric & ~pattern
    assert not missing.any(), (name, np.argwhere(missing)[:5])
    if name in EXACT:
        assert np.array_equal(pattern, numeric), (name, int(pattern.sum()), int(numeric.sum()))
    else:
        assert pattern.sum() <= 3 * numeric.sum()             # upwind / WENO: both wind directions are structural
    plan.close()
