import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, sympy as sp
import _mol_import, mol_b200
from mol_b200 import capi
from mol_b200.interface import Differential, Eq, Interval, MOLFiniteDifference, PDESystem, UpwindScheme, WENOScheme
from mol_b200.lowering import StencilLoweringError
from oracle.discretize import OracleProblem
from ir_interp import IRProgram
seed = int(sys.argv[1]); N = int(sys.argv[2]); emu = len(sys.argv) > 3
rng = np.random.default_rng(seed)
t, x, y, z = sp.symbols("t x y z")
u = sp.Function("u")
def make():
    U = u(t, x, y, z)
    D = {s: Differential(s) for s in (x, y, z)}
    Dt = Differential(t)
    order = int(rng.choice([2, 4]))
    terms = [float(rng.uniform(0.1, 1)) * sum((D[s] ** 2)(U) for s in (x, y, z))]
    a = int(rng.integers(3))
    if a == 1: terms.append(-sum(float(rng.uniform(-1, 1)) * D[s](U) for s in (x, y, z)))
    if a == 2: terms.append(-U * D[x](U) - 0.5 * D[z](U))
    if rng.integers(2): terms.append(U * (1 - U))
    grids = {}
    for s in (x, y, z):
        n = int(rng.integers(8, 12)) if not emu else int(rng.integers(20, 40))
        grids[s] = 1.0 / (n - 1)
    bcs = [Eq(u(0, x, y, z), sp.cos(2 * x) * sp.sin(y + 0.3) * sp.cos(z) + 1.5)]
    per = []
    for s in (x, y, z):
        at = lambda val, s=s: u(t, *[val if q == s else q for q in (x, y, z)])
        others = [q for q in (x, y, z) if q != s]
        if rng.integers(3) == 0:
            bcs.append(Eq(at(0.0), at(1.0))); per.append(True); continue
        per.append(False)
        for end in (0.0, 1.0):
            kind = rng.choice(["dir", "neu", "rob"])
            if kind == "dir": bcs.append(Eq(at(end), sp.exp(-t) * (1.3 + others[0] * others[1])))
            elif kind == "neu": bcs.append(Eq(D[s](at(end)), 0.2 * sp.exp(-t) * others[0]))
            else: bcs.append(Eq(D[s](at(end)) + 1.5 * at(end), sp.cos(t) + others[1]))
    sys_ = PDESystem([Eq(Dt(U), sum(terms))], bcs, [Interval(t, 0.0, 1.0)] + [Interval(s, 0.0, 1.0) for s in (x, y, z)], [t, x, y, z], [U])
    weno = bool(rng.integers(3) == 0)
    return sys_, MOLFiniteDifference(grids, t, approx_order=order, advection_scheme=WENOScheme() if weno else UpwindScheme()), dict(order=order, per=per, weno=weno, terms=str(terms), bcs=[str(b) for b in bcs[1:]])
ok = bad = rej = 0
for k in range(N):
    sys_, disc, info = make()
    try:
        prog = mol_b200.symbolic_discretize(sys_, disc)
    except StencilLoweringError as e:
        rej += 1; print("REJ", str(e)[:100]); continue
    try:
        orc = OracleProblem(sys_, disc)
        uu = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        ref = orc.rhs(uu, 0.37); sc = float(np.max(orc.rhs_termscale(uu, 0.37)))
        if emu:
            from cuda_emu import EmuKernel
            from test_generated_code_cpu import _core_mask
            plan = capi.Plan(prog.text, device=-1)
            got = EmuKernel(plan, prog).rhs([uu], [1.0], 0.37)
            e1 = float(np.max(np.abs(got - ref))) / sc
            e2 = 0.0
            if prog.corebox is not None:
                m = _core_mask(prog)
                g2 = EmuKernel(plan, prog, tiled=True).rhs([uu], [1.0], 0.37)
                e2 = float(np.max(np.abs(g2[m] - ref[m]))) / sc
            plan.close()
            err = max(e1, e2)
        else:
            err = float(np.max(np.abs(IRProgram(prog.text).rhs(uu, 0.37) - ref))) / sc
        if err <= 1e-12: ok += 1
        else: bad += 1; print("MISMATCH", err, info)
    except Exception as e:
        import traceback; traceback.print_exc(); bad += 1; print("EXC", info)
print("ok", ok, "rej", rej, "bad", bad)
