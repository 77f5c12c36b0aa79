import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _mol_import, mol_b200
from mol_b200.lowering import StencilLoweringError
from mol_b200.interface import UpwindScheme
from oracle.discretize import OracleProblem
from ir_interp import IRProgram
from test_random_problems_cpu import random_problem, random_system_2d
rng = np.random.default_rng(int(sys.argv[1])); ok = bad = rej = 0
for k in range(int(sys.argv[2])):
    sys_, disc, what = random_problem(rng) if k % 3 else random_system_2d(rng)
    order = int(rng.choice([2, 3]))
    disc.advection_scheme = UpwindScheme(order)
    try:
        prog = mol_b200.symbolic_discretize(sys_, disc)
    except StencilLoweringError as e:
        rej += 1; print("REJ", str(e)[:110]); continue
    try:
        orc = OracleProblem(sys_, disc)
        u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        ref = orc.rhs(u, 0.37); sc = float(np.max(orc.rhs_termscale(u, 0.37)))
        err = float(np.max(np.abs(IRProgram(prog.text).rhs(u, 0.37) - ref))) / sc
        if err <= 1e-12: ok += 1
        else: bad += 1; print("MISMATCH", err, order, what[:300])
    except (AssertionError, IndexError) as e:
        rej += 1; print("ORACLE-REJ", type(e).__name__, str(e)[:100], order, what[:120])
    except Exception as e:
        import traceback; traceback.print_exc(); bad += 1; print("EXC", order, what[:300])
print("ok", ok, "rej", rej, "bad", bad)
