import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _mol_import, mol_b200
from mol_b200 import capi
from oracle.discretize import OracleProblem
from cuda_emu import EmuKernel
from test_random_problems_cpu import random_problem, random_system_2d
rng = np.random.default_rng(int(sys.argv[1]))
bad = 0
for k in range(int(sys.argv[2])):
    if k % 2:
        sys_, disc, what = random_problem(rng)
    else:
        try:
            sys_, disc, what = random_system_2d(rng, 10, 16)
            prog = mol_b200.symbolic_discretize(sys_, disc)
        except Exception as e:
            continue
    prog = mol_b200.symbolic_discretize(sys_, disc)
    orc = OracleProblem(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    n = orc.nstate
    u = orc.u0 + 0.05 * rng.standard_normal(n); v = rng.standard_normal(n)
    got = EmuKernel(plan, prog, jvp=True).jvp(u, v, 0.37)
    errs = []
    for h in (1e-5, 1e-6):
        want = (orc.rhs(u + h * v, 0.37) - orc.rhs(u - h * v, 0.37)) / (2 * h)
        errs.append(float(np.max(np.abs(got - want)) / max(1.0, float(np.max(np.abs(want))))))
    # unpack too
    full = EmuKernel(plan, prog, unpack=True).unpack(u, 0.37).reshape(len(prog.ilo), -1)
    ref = orc.full_state(u, 0.37)
    ue = max(float(np.max(np.abs(full[w] - np.asarray(ref[w]).reshape(-1, order="F")))) for w in range(len(prog.ilo)))
    if min(errs) > 5e-6 or ue > 1e-11 * max(1.0, float(np.max(np.abs(u)))):
        bad += 1; print("BAD", errs, ue, what[:300])
    plan.close()
print("bad", bad)
