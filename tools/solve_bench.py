"""Diagnostic: adaptive Tsit5 solve of the Brusselator at N^2, time per attempt with the step controller on the device
(queued attempts, csrc/mol_rk.cu solve_queued) and on the host (MOL_RK_QUEUED=0)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import numpy as np
import torch
import mol_b200
from mol_b200 import capi
import problems as examples

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda", 0)
t1 = nsteps * 0.3 * (1.0 / N) ** 2 / 10.0          # ~ nsteps stability-limited steps (alpha = 10)
for mode in ("1", "0", "1", "0"):
    os.environ["MOL_RK_QUEUED"] = mode
    prob = mol_b200.discretize(*examples.brusselator_2d(N, tmax=t1))
    u = torch.from_numpy(prob.u0).to(dev)
    rk = capi.RK(prob.plan, "tsit5", 1e-6, 1e-3)
    st = torch.cuda.current_stream(dev).cuda_stream
    best = None
    for rep in range(3):
        u.copy_(torch.from_numpy(prob.u0))
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        s = rk.solve(u.data_ptr(), 0.0, t1, 0.0, True, stream=st)
        torch.cuda.synchronize()
        el = time.perf_counter() - w0
        att = s.naccept + s.nreject
        best = el if best is None else min(best, el)
    print(f"N={N} queued={mode}: naccept {s.naccept} nreject {s.nreject} nf {s.nf}; best of 3: {best*1e3:.2f} ms = {best/att*1e6:.1f} us per attempt", flush=True)
    rk.close()
