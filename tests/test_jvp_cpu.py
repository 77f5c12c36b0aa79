"""mol_jvp: the generated Jacobian-vector-product kernel -- the stencil program's equations on dual
numbers -- compiled with the host shim and run on the CPU (tests/cuda_emu), against directional derivatives of the
oracle's RHS.  For problems that are affine in u the product is exact (J v = f(v) - f(0)); otherwise a central difference
of the oracle bounds it to truncation error."""
import numpy as np
import pytest

import mol_b200
from mol_b200 import capi
from oracle.discretize import OracleProblem

from cuda_emu import EmuKernel
from test_ir_semantics_cpu import CASES

AFFINE = {"heat_neumann", "heat_robin", "heat_dirichlet_o4", "edge_heat_neumann", "edge_heat_robin_o4", "nu_heat_dirichlet",
          "nu_heat_dirichlet_neumann", "diffusion2d_o4", "heat_dirichlet_o6", "robin_parameter_coefficient",
          "robin_time_dependent_2d", "edge_robin_parameter_coefficient", "beam_two_bcs_at_free_end",
          "mixed_derivative", "mixed_derivative_periodic_y"}


JVP_CASES = ["brusselator", "brusselator_o4", "heat_robin", "heat_dirichlet_o6", "burgers_upwind", "burgers_upwind_nu", "burgers_weno",
             "advection_weno_stretched", "nonlinear_diffusion", "spherical_o4", "burgers2d", "burgers2d_nu", "advection2d_weno",
             "fisher3d_dirichlet_z", "edge_heat_robin_o4", "edge_burgers2d", "nu_heat_dirichlet_neumann", "diffusion2d_o4",
             "robin_parameter_coefficient", "robin_time_dependent_2d", "edge_robin_parameter_coefficient",
             "kdv_three_bcs_per_end", "beam_two_bcs_at_free_end", "mixed_derivative", "mixed_derivative_periodic_y"]


@pytest.mark.parametrize("name", JVP_CASES)
def test_generated_jvp_kernel_matches_directional_derivative(name):
    sys_, disc = CASES[name]()
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    assert plan.cubin("jvp")[:4] == b"\x7fELF"                 # the JVP variant also compiles for sm_100a
    orc = OracleProblem(sys_, disc)
    n = orc.nstate
    rng = np.random.default_rng(21)
    u = orc.u0 + 0.05 * rng.standard_normal(n)
    if name.startswith("nonlinear") or name.startswith("spherical"):
        u = np.abs(u) + 0.1
    v = rng.standard_normal(n)
    t = 0.37
    got = EmuKernel(plan, prog, jvp=True).jvp(u, v, t)
    if name in AFFINE:
        want = orc.rhs(v, t) - orc.rhs(np.zeros(n), t)
        scale = float(np.max(orc.rhs_termscale(v, t)))
        assert np.max(np.abs(got - want)) <= 1e-12 * scale, (name, float(np.max(np.abs(got - want)) / scale))
    else:
        errs = []
        for h in (1e-5, 1e-6):
            want = (orc.rhs(u + h * v, t) - orc.rhs(u - h * v, t)) / (2 * h)
            errs.append(float(np.max(np.abs(got - want)) / max(1.0, float(np.max(np.abs(want))))))
        assert min(errs) <= 2e-6, (name, errs)
    plan.close()
