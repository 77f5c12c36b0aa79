"""Diagnostic: device-resident throughput of J*v (mol_jvp, forward-mode differentiation of the generated equations;
table-driven kernel) beside one RHS evaluation, on the Brusselator, the non-uniform 2-D Burgers problem and 1-D WENO5."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import numpy as np
import torch
import mol_b200
import problems as examples

dev = torch.device("cuda", 0)
cases = {"bruss_2048": lambda: examples.brusselator_2d(2048),
         "burgers2d_nu_2048": lambda: examples.burgers_2d(grid_x=0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, 2049)) / np.tanh(2.0)),
                                                            grid_y=np.linspace(0, 1, 2049) ** 1.3),
         "weno1d_2^22": lambda: examples.advection_1d_periodic(dx=2.0 / (1 << 22), scheme=mol_b200.WENOScheme())}
for name, mk in cases.items():
    prob = mol_b200.discretize(*mk())
    n = prob.plan.state_len
    st = torch.cuda.current_stream(dev).cuda_stream
    u = torch.rand(n, dtype=torch.float64, device=dev) + 0.5
    v = torch.rand(n, dtype=torch.float64, device=dev)
    out = torch.empty_like(u)
    res = {}
    for what in ("rhs", "jvp"):
        call = (lambda: prob.plan.rhs(out.data_ptr(), u.data_ptr(), 0.0, stream=st)) if what == "rhs" else \
               (lambda: prob.plan.jvp(out.data_ptr(), u.data_ptr(), v.data_ptr(), 0.0, stream=st))
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 20
        e0.record()
        for _ in range(K):
            call()
        e1.record(); torch.cuda.synchronize()
        res[what] = e0.elapsed_time(e1) / K * 1e3
    # J*v reads u and v and writes jv: 24 B per unknown
    print(f"{name}: {n} unknowns, rhs {res['rhs']:.1f} us, jvp {res['jvp']:.1f} us = {n * 24 / (res['jvp'] * 1e-6) / 1e9:.0f} GB/s algorithmic "
          f"({n * 24 / (res['jvp'] * 1e-6) / 1e9 / 6546.9:.3f} of the measured copy rate)", flush=True)
