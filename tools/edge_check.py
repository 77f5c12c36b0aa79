"""Diagnostic: edge-aligned grids on the GPU vs the oracle (RHS parity, both kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _mol_import  # noqa
import mol_b200
from mol_b200 import capi, edge_align
import problems as examples
from oracle.discretize import OracleProblem


def edge(sys_, disc):
    return sys_, mol_b200.MOLFiniteDifference(disc.dxs, disc.time, approx_order=disc.approx_order,
                                              advection_scheme=disc.advection_scheme, grid_align=edge_align)


for name, mk in (("heat_neumann", lambda: examples.heat_1d_neumann(dx=0.05)), ("heat_robin_o4", lambda: examples.heat_1d_robin_order4(dx=0.05)),
                 ("burgers2d", lambda: examples.burgers_2d(nx=40, ny=36))):
    sys_, disc = edge(*mk())
    prob = mol_b200.discretize(sys_, disc)
    orc = OracleProblem(sys_, disc)
    u = orc.u0 + 0.05 * np.random.default_rng(7).standard_normal(orc.nstate)
    ref = orc.rhs(u, 0.37)
    scale = float(np.max(orc.rhs_termscale(u, 0.37)))
    for mode in (capi.KERNEL_AUTO, capi.KERNEL_GENERIC):
        prob.plan.set_option("kernel", mode)
        err = float(np.max(np.abs(prob.rhs_host(u, 0.37) - ref)))
        print(name, "mode", mode, "err/termscale", err / scale, "err/max|du|", err / np.max(np.abs(ref)), flush=True)
