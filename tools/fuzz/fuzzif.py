import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, sympy as sp
import _mol_import, mol_b200
from mol_b200.interface import Differential, Eq, Interval, MOLFiniteDifference, PDESystem, UpwindScheme, WENOScheme
from mol_b200.lowering import StencilLoweringError
from oracle.discretize import OracleProblem
from ir_interp import IRProgram
rng = np.random.default_rng(int(sys.argv[1]))
t = sp.Symbol("t")
def make():
    nseg = int(rng.integers(2, 5))
    xs = sp.symbols("x1:%d" % (nseg + 1)); us = [sp.Function("u%d" % (k + 1)) for k in range(nseg)]
    edges = np.concatenate([[0.0], np.cumsum(rng.uniform(0.3, 0.8, nseg))])
    uniform = bool(rng.integers(2))
    weno = bool(rng.integers(2))
    diffusion = uniform and not weno and bool(rng.integers(2))
    dx = 0.02
    if uniform:
        edges = np.round(edges / dx) * dx
    grids = {}
    for k in range(nseg):
        if uniform: grids[xs[k]] = dx
        else:
            n = int(rng.integers(9, 20))
            g = np.sort(np.concatenate([[edges[k], edges[k + 1]], rng.uniform(edges[k], edges[k + 1], n - 2)]))
            if np.diff(g).min() < 1e-3 * (edges[k + 1] - edges[k]): g = np.linspace(edges[k], edges[k + 1], n)
            grids[xs[k]] = g
    eqs, bcs = [], []
    for k in range(nseg):
        U = us[k](t, xs[k]); Dx = Differential(xs[k])
        v = float(rng.uniform(-1.5, 1.5))
        rhs = -v * Dx(U) if rng.integers(2) else -U * Dx(U)
        if diffusion: rhs = rhs + 0.3 * (Dx ** 2)(U)
        if rng.integers(2): rhs = rhs + sp.sin(xs[k]) * U
        eqs.append(Eq(Differential(t)(U), rhs))
        bcs.append(Eq(us[k](0, xs[k]), sp.sin(2 * xs[k]) + 1.2))
    for k in range(nseg - 1):
        a, b = us[k](t, float(edges[k + 1])), us[k + 1](t, float(edges[k + 1]))
        bcs.append(Eq(a, b) if rng.integers(2) else Eq(b, a))
    for end, k, val in ((0, 0, float(edges[0])), (1, nseg - 1, float(edges[-1]))):
        kind = rng.choice(["dir", "neu", "none"]) if not weno else rng.choice(["dir", "neu"])
        Dx = Differential(xs[k])
        if kind == "dir": bcs.append(Eq(us[k](t, val), sp.exp(-t)))
        elif kind == "neu": bcs.append(Eq(Dx(us[k](t, val)), 0.1 * sp.cos(t)))
    dom = [Interval(t, 0.0, 1.0)] + [Interval(xs[k], float(edges[k]), float(edges[k + 1])) for k in range(nseg)]
    sys_ = PDESystem(eqs, bcs, dom, [t] + list(xs), [us[k](t, xs[k]) for k in range(nseg)])
    return sys_, MOLFiniteDifference(grids, t, advection_scheme=WENOScheme() if weno else UpwindScheme()), dict(nseg=nseg, uniform=uniform, weno=weno, diffusion=diffusion, eqs=[str(e) for e in eqs], bcs=[str(b) for b in bcs[nseg:]])
ok = bad = rej = 0
for k in range(int(sys.argv[2])):
    sys_, disc, info = make()
    try:
        prog = mol_b200.symbolic_discretize(sys_, disc)
    except StencilLoweringError as e:
        rej += 1; print("REJ", str(e)[:110], info["uniform"], info["weno"]); continue
    try:
        orc = OracleProblem(sys_, disc)
        uu = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        ref = orc.rhs(uu, 0.37); sc = float(np.max(orc.rhs_termscale(uu, 0.37)))
        err = float(np.max(np.abs(IRProgram(prog.text).rhs(uu, 0.37) - ref))) / sc
        if err <= 1e-12: ok += 1
        else: bad += 1; print("MISMATCH", err, info)
    except Exception as e:
        import traceback; traceback.print_exc(); bad += 1; print("EXC", info)
print("ok", ok, "rej", rej, "bad", bad)
