import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _mol_import, mol_b200
from mol_b200 import capi
from mol_b200.lowering import StencilLoweringError
from oracle.discretize import OracleProblem
from cuda_emu import EmuKernel
from test_random_problems_cpu import random_system_2d, random_problem
from test_generated_code_cpu import _core_mask
rng = np.random.default_rng(int(sys.argv[1])); bad = ok = 0
for k in range(int(sys.argv[2])):
    try:
        sys_, disc, what = random_system_2d(rng, 40, 70) if k % 2 == 0 else random_problem(rng)
        prog = mol_b200.symbolic_discretize(sys_, disc)
    except StencilLoweringError:
        continue
    orc = OracleProblem(sys_, disc); n = orc.nstate
    plan = capi.Plan(prog.text, device=-1)
    u = orc.u0 + 0.05 * rng.standard_normal(n)
    k1, k2 = 0.3 * rng.standard_normal(n), 0.3 * rng.standard_normal(n)
    coefs = [1.0, 0.01, -0.02]; uin = u + 0.01 * k1 - 0.02 * k2
    ref = orc.rhs(uin, 0.37); sc = float(np.max(orc.rhs_termscale(uin, 0.37)))
    e = float(np.max(np.abs(EmuKernel(plan, prog, nin=3).rhs([u, k1, k2], coefs, 0.37) - ref))) / sc
    e2 = 0.0
    if prog.corebox is not None:
        m = _core_mask(prog)
        for staging in ("coop",):
            got = EmuKernel(plan, prog, nin=3, tiled=True).rhs([u, k1, k2], coefs, 0.37)
            e2 = max(e2, float(np.max(np.abs(got[m] - ref[m]))) / sc)
    plan.close()
    if max(e, e2) <= 1e-12: ok += 1
    else: bad += 1; print("BAD", e, e2, what[:200])
print("ok", ok, "bad", bad)
