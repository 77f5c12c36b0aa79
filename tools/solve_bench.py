"""Diagnostic (SURVEY §8d "CPU baseline beside it", item 2): one adaptive Tsit5 solve of the 512^2 Brusselator on the
GPU (mol_rk_solve: fused stages, device-resident state) beside the CPU restatement (oracle Tsit5 loop in NumPy around
oracle/bruss_ref.c with all host threads), same tolerances, same tspan; also checks the two final states agree.
usage: python tools/solve_bench.py [N=512] [tend=2e-4]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import torch
import mol_b200
import problems as examples
from oracle import cref
from oracle.rk import solve_tsit5

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tend = float(sys.argv[2]) if len(sys.argv) > 2 else 2e-4
abstol, reltol = 1e-6, 1e-3
sys_, disc = examples.brusselator_2d(N, tmax=tend)
prob = mol_b200.discretize(sys_, disc)
u0 = prob.u0
mol_b200.solve(prob, mol_b200.Tsit5(), abstol=abstol, reltol=reltol)          # warm-up (NVRTC variants, allocations)
torch.cuda.synchronize()
t0 = time.perf_counter()
sol = mol_b200.solve(prob, mol_b200.Tsit5(), abstol=abstol, reltol=reltol)
torch.cuda.synchronize()
gpu_s = time.perf_counter() - t0

xg, yg = prob.program.axes[0].x, prob.program.axes[1].x
nth = cref.lib().bruss_ref_max_threads()
f = lambda u, t: cref.bruss_rhs(u, xg, yg, N, t, nthreads=nth)
t0 = time.perf_counter()
ts, us, st = solve_tsit5(f, u0, (0.0, tend), abstol=abstol, reltol=reltol)
cpu_s = time.perf_counter() - t0
err = float(np.max(np.abs(sol.u[-1] - us[-1]) / (abstol + reltol * np.abs(us[-1]))))
print(json.dumps({"workload": f"brusselator2d_{N}x{N} Tsit5 adaptive 0..{tend:g} abstol={abstol:g} reltol={reltol:g}",
                  "gpu_seconds": gpu_s, "gpu_steps": sol.stats["naccept"] + sol.stats["nreject"], "gpu_nf": sol.stats["nf"],
                  "gpu_us_per_step": 1e6 * gpu_s / max(1, sol.stats["naccept"] + sol.stats["nreject"]),
                  "cpu_seconds": cpu_s, "cpu_threads": nth, "cpu_steps": st["naccept"] + st["nreject"], "cpu_nf": st["nf"],
                  "cpu_kind": "port (NumPy Tsit5 loop around oracle/bruss_ref.c, OpenMP)",
                  "speedup": cpu_s / gpu_s, "max_scaled_final_state_difference": err}))
