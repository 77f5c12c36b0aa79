"""The stencil program the lowering emits, executed by the NumPy IR interpreter (tests/ir_interp.py), against the
oracle: checks tables, ghost rules and equations of every scheme family on a machine without a GPU."""
import numpy as np
import pytest

import mol_b200
from mol_b200 import edge_align
import problems as examples
from oracle.discretize import OracleProblem

from ir_interp import IRProgram


def _edge(sys_, disc):
    return sys_, mol_b200.MOLFiniteDifference(disc.dxs, disc.time, approx_order=disc.approx_order,
                                              advection_scheme=disc.advection_scheme, grid_align=edge_align)


def _order(sd, p):
    sd[1].approx_order = p
    return sd


CASES = {
    "brusselator": lambda: examples.brusselator_2d(8),
    "brusselator_o4": lambda: examples.brusselator_2d(10, approx_order=4),
    "heat_dirichlet_o4": lambda: examples.heat_1d_dirichlet(dx=0.05, approx_order=4),
    "heat_neumann": lambda: examples.heat_1d_neumann(dx=0.05),
    "heat_robin": lambda: examples.heat_1d_robin(dx=0.05),
    "burgers_upwind": lambda: examples.burgers_1d(dx=0.05),
    "burgers_upwind_nu": lambda: examples.burgers_1d(grid=examples.stretched_grid(0, 1, 31, 0.03)),
    "burgers_weno": lambda: examples.burgers_1d(dx=0.05, scheme=mol_b200.WENOScheme()),
    "burgers_weno_nu": lambda: examples.burgers_1d(grid=examples.stretched_grid(0, 1, 31, 0.03), scheme=mol_b200.WENOScheme()),
    "advection_weno_periodic": lambda: examples.advection_1d_periodic(dx=0.05, scheme=mol_b200.WENOScheme()),
    "advection_weno_stretched": lambda: examples.advection_1d_periodic(dx=examples.stretched_grid(0, 2, 32), scheme=mol_b200.WENOScheme()),
    "nonlinear_diffusion": lambda: examples.nonlinear_diffusion_1d(dx=0.05),
    "spherical": lambda: examples.spherical_diffusion_1d(dr=0.1),
    "spherical_o4": lambda: examples.spherical_diffusion_order4(dr=0.1),
    "burgers2d": lambda: examples.burgers_2d(nx=10, ny=9),
    "burgers2d_nu": lambda: examples.burgers_2d(grid_x=0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, 11)) / np.tanh(2.0)),
                                                grid_y=np.linspace(0, 1, 10) ** 1.3),
    "advection2d_weno": lambda: examples.advection_2d_periodic(8, scheme=mol_b200.WENOScheme()),
    "fisher3d_dirichlet_z": lambda: examples.diffusion_reaction_3d(n=6, periodic=False),
    # higher orders: UpwindScheme(2), UpwindScheme(3), approx_order = 6
    "burgers_upwind_o2": lambda: examples.burgers_1d(dx=0.05, scheme=mol_b200.UpwindScheme(2)),
    "burgers_upwind_o3": lambda: examples.burgers_1d(dx=0.05, scheme=mol_b200.UpwindScheme(3)),
    "advection_periodic_upwind_o2": lambda: examples.advection_1d_periodic(dx=0.05, scheme=mol_b200.UpwindScheme(2)),
    "heat_dirichlet_o6": lambda: examples.heat_1d_dirichlet(dx=0.05, approx_order=6),
    "brusselator_o6": lambda: examples.brusselator_2d(12, approx_order=6),
    # the reference's non-uniform diffusion tests (jittered nodes, orders 2 and 4, Dirichlet + Neumann) and 2-D diffusion
    "nu_heat_dirichlet": lambda: examples.heat_1d_dirichlet_pi(examples.jittered_grid(0.0, float(np.pi), 30)),
    "nu_heat_dirichlet_o4": lambda: examples.heat_1d_dirichlet_pi(examples.jittered_grid(0.0, float(np.pi), 30), approx_order=4),
    "nu_heat_dirichlet_neumann": lambda: examples.heat_1d_dirichlet_neumann_pi(examples.jittered_grid(0.0, float(np.pi), 30)),
    "diffusion2d_o4": lambda: examples.diffusion_2d_dirichlet(),
    # ghost rules with expression coefficients (a parameter / a time- and position-dependent Robin coefficient)
    "robin_parameter_coefficient": lambda: examples.advection_diffusion_robin_param(dx=0.05),
    "robin_time_dependent_2d": lambda: examples.heat_2d_robin_time_dependent(nx=12, ny=10),
    "edge_robin_parameter_coefficient": lambda: _edge(*examples.advection_diffusion_robin_param(dx=0.05)),
    "three_species": lambda: examples.three_species_2d(12, 10),
    # test/Diffusion/MOL_1D_Linear_Diffusion.jl Tests 06 (time-dependent Robin coefficients, order 6), 10 (two variables,
    # opposite Dirichlet / Neumann ends), 11 (parameter diffusivities + reaction)
    "nonlinear_diffusion_2d": lambda: examples.nonlinear_diffusion_2d(),
    "spherical_outer_coefficient": lambda: examples.spherical_diffusion_coefficient4(dr=0.05),
    "heat_parameter_diffusivity": lambda: examples.heat_parameter_diffusivity(),
    "diffusion_variable_coefficient": lambda: examples.diffusion_variable_coefficient(),
    "heat_robin_time_dependent_o6": lambda: examples.heat_1d_robin_time_dependent(dx=0.05),
    "two_variables_mixed_bcs": lambda: examples.diffusion_two_variables_mixed_bcs(l=30),
    "reaction_diffusion_parameters": lambda: examples.reaction_diffusion_parameters(),
    # mixed derivative Dx Dy u (2nd_order_mixed_deriv.jl): corner nodes read as 0, periodic taps wrapped first
    "mixed_derivative": lambda: examples.anisotropic_diffusion_2d(12, 10),
    "mixed_derivative_periodic_y": lambda: examples.anisotropic_diffusion_2d(12, 10, periodic_y=True),
    # several boundary conditions at one end (test/Higher_Order/MOL_1D_HigherOrder.jl:51-152): the clipped nodes solve an
    # affine system; `v ~ Dt(u)` (coefficient -1 on the time derivative)
    "kdv_three_bcs_per_end": lambda: examples.kdv_soliton(),
    "beam_two_bcs_at_free_end": lambda: examples.beam_with_velocity(),
    # nonlinear Laplacian on a jittered grid (test/Nonlinear_Diffusion_NU/...:135-262), orders 2 and 4
    "nonlinear_diffusion_nu": lambda: examples.nonlinear_diffusion_travelling(dx=examples.jittered_grid(0.0, 2.0, 41, 1e-3)),
    "nonlinear_diffusion_nu_o4": lambda: _order(examples.nonlinear_diffusion_travelling(dx=examples.jittered_grid(0.0, 2.0, 41, 1e-3)), 4),
    # variables on different domains joined by interface boundary conditions (interface_boundary.jl:79-153): one chart
    # axis in the stencil program, per-variable grids in the oracle (oracle/interface1d.py)
    "two_independent_domains_o4": lambda: examples.diffusion_two_independent_domains(l=20, approx_order=4),
    "pde_with_ode": lambda: examples.diffusion_with_ode(l=20),
    "pde_driven_by_ode": lambda: examples.diffusion_driven_by_ode(l=20),
    "iface_diffusion": lambda: examples.diffusion_two_domains(),
    "iface_diffusion_o4": lambda: examples.diffusion_two_domains(l=14, approx_order=4),
    "iface_upwind_nu": lambda: examples.advection_two_domains(),
    "iface_upwind_nu_neg": lambda: examples.advection_two_domains(v=-1.0),
    "iface_upwind_nu_opposed": lambda: examples.advection_two_domains(v=1.0, v2=-0.5),
    "iface_upwind_uniform": lambda: examples.advection_two_domains(x1grid=0.02, x2grid=0.02),
    "iface_upwind_chain4": lambda: examples.advection_chained_domains(),
    # uniform WENO5 with a Neumann outflow end: the derivative condition and the extrapolation pad are solved together
    "iface_weno_uniform_neumann": lambda: examples.advection_two_domains(x1grid=0.02, x2grid=0.02, scheme=mol_b200.WENOScheme()),
    "weno_uniform_neumann_outflow": lambda: examples.advection_inflow_nu([0.0, 1.0], v=0.8, scheme=mol_b200.WENOScheme(), dx=0.025),
    "iface_weno_nu": lambda: examples.advection_two_domains(scheme=mol_b200.WENOScheme()),
    "iface_weno_nu_neg": lambda: examples.advection_two_domains(scheme=mol_b200.WENOScheme(), v=-1.0),
    "iface_weno_chain4": lambda: examples.advection_chained_domains(scheme=mol_b200.WENOScheme()),
    # periodic dimensions on an edge-aligned grid (node 1 identified with node n like on any grid)
    "edge_advection2d_periodic": lambda: _edge(*examples.advection_2d_periodic(14, nu=0.01)),
    "edge_heat_neumann": lambda: _edge(*examples.heat_1d_neumann(dx=0.05)),
    "edge_heat_robin_o4": lambda: _edge(*examples.heat_1d_robin_order4(dx=0.05)),
    "edge_burgers2d": lambda: _edge(*examples.burgers_2d(nx=10, ny=9)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_stencil_program_semantics_match_oracle(name):
    sys_, disc = CASES[name]()
    prog = mol_b200.symbolic_discretize(sys_, disc)
    orc = OracleProblem(sys_, disc)
    ir = IRProgram(prog.text)
    assert ir.nstate == orc.nstate == prog.nstate
    rng = np.random.default_rng(9)
    u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
    if name.startswith("nonlinear") or name.startswith("spherical"):
        u = np.abs(u) + 0.1
    for t in (0.0, 0.37):
        ref = orc.rhs(u, t)
        got = ir.rhs(u, t)
        scale = float(np.max(orc.rhs_termscale(u, t)))
        err = float(np.max(np.abs(got - ref)))
        assert err <= 1e-13 * scale, (name, t, err / scale)
        assert err <= 1e-12 * np.max(np.abs(ref)), (name, t, err / np.max(np.abs(ref)))
