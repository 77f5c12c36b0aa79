"""Diagnostic: queued vs host-driven adaptive Tsit5, with and without save points (final states, counters)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import numpy as np
import torch
import mol_b200
from mol_b200 import capi
import problems as examples

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t1 = 2e-3
res = {}
for mode, batch in (("0", "8"), ("1", "8"), ("1", "1"), ("1", "3")):
    os.environ["MOL_RK_QUEUED"] = mode
    os.environ["MOL_RK_QUEUED_BATCH"] = batch
    prob = mol_b200.discretize(*examples.brusselator_2d(N, tmax=t1))
    n = prob.plan.state_len
    st = torch.cuda.current_stream(dev).cuda_stream
    for saves in ([], [0.0, 3.3e-4, 1.0e-3, 1.9e-3, 2e-3]):
        rk = capi.RK(prob.plan, "tsit5", 1e-6, 1e-3)
        for rep in range(2):
            u = torch.from_numpy(prob.u0).to(dev)
            save = torch.zeros((max(1, len(saves)), n), dtype=torch.float64, device=dev)
            s = rk.solve(u.data_ptr(), 0.0, t1, 0.0, True, np.array(saves) if saves else None, save.data_ptr() if saves else 0, stream=st)
            torch.cuda.synchronize()
            key = (mode, batch, bool(saves), rep)
            res[key] = (u.cpu().numpy(), save.cpu().numpy(), s)
            print(key, "naccept", s.naccept, "nreject", s.nreject, "nf", s.nf, "t_final", s.t_final, "dt_last", s.dt_last, "ret", s.retcode, flush=True)
        rk.close()
ref = res[("0", "8", False, 0)][0]
for k, (u, sv, s) in res.items():
    line = f"{k}: |u - host_nosave| = {np.max(np.abs(u - ref)):.3e}"
    if k[2]:
        line += f"; |save[-1] - u| = {np.max(np.abs(sv[-1] - u)):.3e}; |save[2] - host_save[2]| = {np.max(np.abs(sv[2] - res[('0', '8', True, 0)][1][2])):.3e}"
    print(line)
