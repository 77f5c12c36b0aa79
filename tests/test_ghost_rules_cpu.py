"""CPU cross-check of the lowering's ghost rules against the oracle's boundary reconstruction.

The two are independent restatements of generate_bc_eqs.jl: the lowering solves each boundary equation symbolically for
the boundary node, u[node] = G(t, x_perp) + sum_k a_k u[tap_k] (what the CUDA kernels evaluate on the fly); the oracle
fills the boundary nodes of full-grid arrays numerically (OracleProblem.full_state).  Evaluating the rules in NumPy on
seeded states must reproduce the oracle's boundary nodes -- for centre-aligned and edge-aligned grids
(generate_bc_eqs.jl:238-311 / :79-161), Dirichlet, Neumann, Robin and the extrapolation pads of uniform WENO."""
import numpy as np
import pytest
import sympy as sp

import mol_b200
from mol_b200 import edge_align
import problems as examples
from mol_b200.lowering import Lowering
from oracle.discretize import OracleProblem


def edge(mk):
    def f():
        sys_, disc = mk()
        d2 = mol_b200.MOLFiniteDifference(disc.dxs, disc.time, approx_order=disc.approx_order,
                                          advection_scheme=disc.advection_scheme, grid_align=edge_align)
        return sys_, d2
    return f


CASES = {
    "heat_dirichlet": lambda: examples.heat_1d_dirichlet(dx=0.05),
    "heat_dirichlet_o4": lambda: examples.heat_1d_dirichlet(dx=0.05, approx_order=4),
    "heat_neumann": lambda: examples.heat_1d_neumann(dx=0.05),
    "heat_robin": lambda: examples.heat_1d_robin(dx=0.05),
    "heat_robin_o4": lambda: examples.heat_1d_robin_order4(dx=0.05),
    "burgers_weno_pads": lambda: examples.burgers_1d(dx=0.05, scheme=mol_b200.WENOScheme()),
    "burgers_upwind_nu": lambda: examples.burgers_1d(grid=examples.stretched_grid(0, 1, 31, 0.03)),
    "burgers2d": lambda: examples.burgers_2d(nx=12, ny=10),
    "fisher3d_dirichlet_z": lambda: examples.diffusion_reaction_3d(n=8, periodic=False),
    "robin_parameter_coefficient": lambda: examples.advection_diffusion_robin_param(dx=0.05),
    "robin_time_dependent_2d": lambda: examples.heat_2d_robin_time_dependent(nx=12, ny=10),
}
for k in ("heat_dirichlet", "heat_neumann", "heat_robin", "heat_robin_o4", "burgers2d", "burgers_upwind_nu",
          "robin_parameter_coefficient"):
    CASES[k + "_edge"] = edge(CASES[k])


@pytest.mark.parametrize("name", sorted(CASES))
def test_ghost_rules_reproduce_oracle_boundary_nodes(name):
    sys_, disc = CASES[name]()
    L = Lowering(sys_, disc)
    prog = L.lower()
    orc = OracleProblem(sys_, disc)
    assert prog.nstate == orc.nstate
    for j in range(L.nd):
        np.testing.assert_allclose(L.axes[j].x, orc.grid[j], rtol=0, atol=1e-15)
    np.testing.assert_allclose(prog.u0, orc.u0, rtol=1e-15, atol=1e-16)
    plan = mol_b200.capi.Plan(prog.text, device=-1)          # the generated ghost rules / equations compile for sm_100a
    assert plan.state_len == orc.nstate
    plan.close()
    rules = L._ghosts()
    assert rules, "no boundary rules?"
    rng = np.random.default_rng(5)
    u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
    for t in (0.0, 0.37):
        full = orc.full_state(u, t)
        for (v, j, node), (G, taps) in rules.items():
            # interior extents in the other dimensions, the ghost node in dimension j
            idx = [slice(L.ilo[v][d] - 1, L.ihi[v][d]) for d in range(L.nd)]
            shape = [L.ihi[v][d] - L.ilo[v][d] + 1 for d in range(L.nd)]
            shape[j] = 1
            coords = []
            for d in range(L.nd):
                sh = [1] * L.nd
                g = L.axes[d].x[idx[d]] if d != j else np.zeros(1)
                sh[d] = len(g)
                coords.append(g.reshape(sh))
            f = sp.lambdify(list(L.xs) + [L.t] + list(L.params), G, "numpy")
            val = np.broadcast_to(np.asarray(f(*coords, t, *L.pvals), dtype=float), shape).copy()
            for (w_, tp), a in taps.items():
                sl = list(idx)
                sl[j] = slice(tp - 1, tp)
                if not isinstance(a, float):            # expression coefficient (parameters, t, boundary coordinates)
                    fa = sp.lambdify(list(L.xs) + [L.t] + list(L.params), a, "numpy")
                    a = np.broadcast_to(np.asarray(fa(*coords, t, *L.pvals), dtype=float), shape)
                val += a * np.asarray(full[w_])[tuple(sl)]
            sl = list(idx)
            sl[j] = slice(node - 1, node)
            want = np.asarray(full[v])[tuple(sl)]
            scale = max(1.0, float(np.max(np.abs(want))))
            assert np.max(np.abs(val - want)) <= 1e-12 * scale, (name, t, v, j, node, float(np.max(np.abs(val - want))))
