"""Diagnostic: fused explicit-RK step throughput on the Brusselator (device-resident state)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import torch
import mol_b200
from mol_b200 import capi
import problems as examples

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
algs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tsit5", "ssprk33", "euler"]
dev = torch.device("cuda", 0)
prob = mol_b200.discretize(*examples.brusselator_2d(N))
n = prob.plan.state_len
u = torch.rand(n, dtype=torch.float64, device=dev) * 3
st = torch.cuda.current_stream(dev).cuda_stream
for alg in algs:
    rk = capi.RK(prob.plan, alg, 1e-6, 1e-3)
    t, dt = 0.0, 1e-9
    for _ in range(3):
        t, _, _ = rk.step(u.data_ptr(), t, dt, adaptive=False, stream=st)
    torch.cuda.synchronize()
    K = 30
    l0 = prob.plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        t, _, s = rk.step(u.data_ptr(), t, dt, adaptive=False, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    nf = {"tsit5": 6, "ssprk33": 3, "euler": 1, "rk4": 4}[alg]
    passes = {"tsit5": 35, "ssprk33": 3 * 2 + (2 + 3) + 5, "euler": 2 + 3, "rk4": 0}[alg]
    print(f"{alg}: {ms*1e3:.1f} us/step, {nf} RHS/step -> {N*N*nf/(ms*1e-3):.3e} RHS updates/s, {N*N/(ms*1e-3):.3e} cell-steps/s, "
          f"launches/step {(prob.plan.launch_count()-l0)/K:.1f}", flush=True)
    rk.close()
