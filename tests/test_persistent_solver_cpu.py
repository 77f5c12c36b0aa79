"""The persistent single-CTA solver kernel (kernels/mol_generic.cuh, MOL_KERNEL_SOLVE) on the CPU emulator: a whole
solve -- stages, embedded error norm, PI controller, dense-output saves -- in one kernel, against the oracle's
integrators (oracle/rk.py) on the oracle's RHS: same accepted / rejected step sequence, same saved states."""
import numpy as np
import pytest

import _mol_import  # noqa: F401
import mol_b200
from mol_b200 import capi
import problems as examples
from oracle.discretize import OracleProblem
from oracle.rk import solve_fixed, solve_tsit5
from cuda_emu import EmuKernel


def _setup(mk):
    sys_, disc = mk()
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    return OracleProblem(sys_, disc), EmuKernel(plan, prog, solve=True), plan


def test_config1_heat_adaptive_tsit5_with_saveat():
    orc, emu, plan = _setup(lambda: examples.heat_1d_dirichlet(dx=0.02))
    sv = [0.0, 0.0137, 0.2, 0.731, 1.0]
    ts, us, st = solve_tsit5(orc.rhs, orc.u0, (0.0, 1.0), abstol=1e-8, reltol=1e-8, saveat=sv)
    u1, saved, stats = emu.solve(orc.u0, "tsit5", 0.0, 1.0, adaptive=True, abstol=1e-8, reltol=1e-8, saveat=sv)
    assert stats["retcode"] == 0 and stats["nsaved"] == len(sv)
    assert (stats["naccept"], stats["nreject"]) == (st["naccept"], st["nreject"])
    assert stats["nf"] == st["nf"] - 1          # f(u0) of the starting-step heuristic is reused as k1
    for k in range(len(sv)):
        np.testing.assert_allclose(saved[k], us[k], rtol=0, atol=1e-11)
    np.testing.assert_allclose(u1, us[-1], rtol=0, atol=1e-11)
    plan.close()


def test_brusselator_adaptive_default_tolerances_step_sequence():
    orc, emu, plan = _setup(lambda: examples.brusselator_2d(8, tmax=0.02))
    ts, us, st = solve_tsit5(orc.rhs, orc.u0, (0.0, 0.02))
    u1, _, stats = emu.solve(orc.u0, "tsit5", 0.0, 0.02)
    assert stats["retcode"] == 0 and (stats["naccept"], stats["nreject"]) == (st["naccept"], st["nreject"])
    np.testing.assert_allclose(u1, us[-1], rtol=1e-9, atol=1e-10)
    plan.close()


@pytest.mark.parametrize("alg", ["euler", "ssprk33", "rk4", "tsit5"])
def test_fixed_step_methods_with_misaligned_saveat(alg):
    orc, emu, plan = _setup(lambda: examples.advection_1d_periodic(dx=0.05, scheme=mol_b200.WENOScheme(), tmax=0.1))
    dt, sv = 0.0071, [0.0, 0.01, 0.05, 0.0999, 0.1]
    ts, us = solve_fixed(orc.rhs, orc.u0, (0.0, 0.1), dt, alg, saveat=sv)
    u1, saved, stats = emu.solve(orc.u0, alg, 0.0, 0.1, dt0=dt, adaptive=False, saveat=sv)
    assert stats["retcode"] == 0 and stats["naccept"] == 15 and stats["nsaved"] == len(sv)
    for k in range(len(sv)):
        np.testing.assert_allclose(saved[k], us[k], rtol=0, atol=1e-11)
    plan.close()
