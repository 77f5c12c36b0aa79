"""Diagnostic: BASELINE config 1 (1-D heat, Dirichlet, 101 nodes, Tsit5, saveat = 0.2, docs/src/tutorials/heat.md) and
two other small problems through mol_rk_solve: the persistent single-CTA solver kernel (one launch per solve) against
the host-driven loop (MOL_RK_PERSISTENT=0: seven launches + one read-back per step).
usage: python tools/config1_bench.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import torch
import mol_b200
import problems as examples

cases = {"heat1d_101_tsit5_default_tol": (lambda: examples.heat_1d_dirichlet(dx=0.01), dict(saveat=0.2)),
         "heat1d_101_tsit5_1e-8": (lambda: examples.heat_1d_dirichlet(dx=0.01), dict(saveat=0.2, abstol=1e-8, reltol=1e-8)),
         "bruss_32x32_tsit5_t0.05": (lambda: examples.brusselator_2d(32, tmax=0.05), dict()),
         "weno1d_128_ssprk33": (lambda: examples.advection_1d_periodic(dx=2.0 / 128, scheme=mol_b200.WENOScheme(), tmax=1.0),
                                dict(alg=mol_b200.SSPRK33(), dt=0.4 * 2.0 / 128, adaptive=False))}
out = {}
for name, (mk, kw) in cases.items():
    kw = dict(kw)
    alg = kw.pop("alg", mol_b200.Tsit5())
    res = {}
    for mode in ("1", "0"):
        os.environ["MOL_RK_PERSISTENT"] = mode
        prob = mol_b200.discretize(*mk())
        mol_b200.solve(prob, alg, **kw)                      # warm-up: NVRTC variants, allocations
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            sol = mol_b200.solve(prob, alg, **kw)
        torch.cuda.synchronize()
        s = (time.perf_counter() - t0) / reps
        steps = sol.stats["naccept"] + sol.stats["nreject"]
        res["persistent" if mode == "1" else "host_loop"] = {"solve_ms": 1e3 * s, "steps": steps, "us_per_step": 1e6 * s / max(1, steps),
                                                             "nf": sol.stats["nf"], "final": sol.u[-1]}
    d = float(np.max(np.abs(res["persistent"]["final"] - res["host_loop"]["final"])))
    for r in res.values():
        r.pop("final")
    res["max_abs_difference_of_final_states"] = d
    res["speedup"] = res["host_loop"]["solve_ms"] / res["persistent"]["solve_ms"]
    out[name] = res
print(json.dumps(out))
