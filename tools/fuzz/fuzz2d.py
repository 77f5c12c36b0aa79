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
t, x, y = sp.symbols("t x y")
u, v = sp.Function("u"), sp.Function("v")
def make():
    order = int(rng.choice([2, 4]))
    nvar = int(rng.integers(1, 3))
    fs = [u, v][:nvar]
    Dt, Dx, Dy = Differential(t), Differential(x), Differential(y)
    grids = {}
    ns = {}
    for s_ in (x, y):
        n = int(rng.integers(14, 26)) if not emu else int(rng.integers(40, 80))
        ns[s_] = n
        if rng.integers(3) == 0:
            g = np.linspace(0, 1, n) ** float(rng.uniform(1.0, 1.4)); g[-1] = 1.0
            grids[s_] = g
        else:
            grids[s_] = 1.0 / (n - 1)
    weno = bool(rng.integers(3) == 0)
    per = {x: bool(rng.integers(4) == 0), y: bool(rng.integers(4) == 0)}
    eqs, bcs = [], []
    for k, f in enumerate(fs):
        F = f(t, x, y)
        terms = [float(rng.uniform(0.1, 1)) * ((Dx ** 2)(F) + (Dy ** 2)(F))] if rng.integers(4) else []
        a = int(rng.integers(3))
        if a == 1: terms.append(-float(rng.uniform(-1, 1)) * Dx(F) - float(rng.uniform(-1, 1)) * Dy(F))
        if a == 2: terms.append(-fs[0](t, x, y) * Dx(F) - fs[-1](t, x, y) * Dy(F))
        if rng.integers(2): terms.append(F * (1 - fs[-1](t, x, y)) + sp.sin(x + y) * sp.exp(-t))
        if not terms: terms.append((Dx ** 2)(F))
        eqs.append(Eq(Dt(F), sum(terms)))
        bcs.append(Eq(f(0, x, y), sp.cos(2 * x + k) * sp.sin(y + 0.3) + 1.5))
        for s_, D_ in ((x, Dx), (y, Dy)):
            mk = (lambda val: f(t, val, y)) if s_ == x else (lambda val: f(t, x, val))
            if per[s_]:
                bcs.append(Eq(mk(0.0), mk(1.0))); continue
            for end in (0.0, 1.0):
                kind = rng.choice(["dir", "neu", "rob"])
                other = y if s_ == x else x
                if kind == "dir": bcs.append(Eq(mk(end), sp.exp(-t) * (1.3 + other)))
                elif kind == "neu": bcs.append(Eq(D_(mk(end)), 0.2 * sp.exp(-t) * other))
                else: bcs.append(Eq(D_(mk(end)) + float(rng.uniform(0.5, 2)) * mk(end), sp.cos(t) + other))
    sys_ = PDESystem(eqs, bcs, [Interval(t, 0.0, 1.0), Interval(x, 0.0, 1.0), Interval(y, 0.0, 1.0)], [t, x, y], [f(t, x, y) for f in fs])
    disc = MOLFiniteDifference(grids, t, approx_order=order, advection_scheme=WENOScheme() if weno else UpwindScheme())
    info = dict(order=order, nvar=nvar, weno=weno, per=[per[x], per[y]], nu=[np.ndim(grids[x]) > 0, np.ndim(grids[y]) > 0], n=[ns[x], ns[y]], eqs=[str(e) for e in eqs], bcs=[str(b) for b in bcs if not str(b).startswith("u(0") and not str(b).startswith("v(0")])
    return sys_, disc, info
ok = rej = bad = 0
for k in range(N):
    sys_, disc, info = make()
    try:
        prog = mol_b200.symbolic_discretize(sys_, disc)
    except StencilLoweringError as e:
        rej += 1; print("REJ", str(e)[:120], {q: info[q] for q in ("order", "weno", "per", "nu")}); continue
    try:
        orc = OracleProblem(sys_, disc)
        uu = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        if emu:
            from cuda_emu import EmuKernel
            plan = capi.Plan(prog.text, device=-1)
            runs = [("generic", EmuKernel(plan, prog))]
            mask = None
            if prog.corebox is not None:
                from test_generated_code_cpu import _core_mask
                mask = _core_mask(prog)
                runs.append(("tiled", EmuKernel(plan, prog, tiled=True)))
        else:
            runs = [("ir", IRProgram(prog.text))]
        good = True
        for tt in (0.0, 0.37):
            ref = orc.rhs(uu, tt)
            sc = float(np.max(orc.rhs_termscale(uu, tt)))
            for name, kern in runs:
                got = kern.rhs(uu, tt) if name == "ir" else kern.rhs([uu], [1.0], tt)
                d = np.abs(ref - got)
                if name == "tiled": d = d[mask]
                err = float(np.max(d)) / sc
                if not err <= 1e-12:
                    good = False; print("MISMATCH", name, err, info)
        ok += good; bad += (not good)
        if emu: plan.close()
    except Exception as e:
        import traceback; traceback.print_exc()
        bad += 1; print("EXC", type(e).__name__, str(e)[:300], info)
print("ok", ok, "rejected", rej, "bad", bad)
