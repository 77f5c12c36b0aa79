"""Diagnostic: device-resident RHS throughput of any example problem (single GPU or torchrun slabs)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import torch
import mol_b200
import problems as examples
from mol_b200.distributed import SlabRunner

case = sys.argv[1] if len(sys.argv) > 1 else "fisher3d"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
weak = (sys.argv[3] != "strong") if len(sys.argv) > 3 else True
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
nz = int(os.environ.get("NZ", "0")) or None
mk = {"fisher3d": lambda: examples.diffusion_reaction_3d(n=n, periodic=True, nz=nz),
      "fisher3d_dirichlet": lambda: examples.diffusion_reaction_3d(n=n, periodic=False),
      "bruss": lambda: examples.brusselator_2d(n),
      "burgers2d": lambda: examples.burgers_2d(nx=n, ny=n),
      # config 3: upwind Burgers, Neumann + Robin + Dirichlet, tanh-stretched non-uniform grid in x, power-law grid in y
      "burgers2d_nu": lambda: examples.burgers_2d(grid_x=0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, n + 1)) / np.tanh(2.0)),
                                                  grid_y=np.linspace(0, 1, n + 1) ** 1.3),
      # config 4: WENO5 convection, 1-D (benchmark/weno/problems.jl) and 2-D
      "weno1d": lambda: examples.advection_1d_periodic(dx=2.0 / n, scheme=mol_b200.WENOScheme()),
      "weno1d_burgers": lambda: examples.weno_burgers_periodic(dx=2.0 / n),
      "weno1d_nu": lambda: examples.advection_1d_periodic(dx=examples.stretched_grid(0, 2, n + 1), scheme=mol_b200.WENOScheme()),
      "weno2d": lambda: examples.advection_2d_periodic(n, scheme=mol_b200.WENOScheme()),
      "weno2d_nu": lambda: examples.advection_2d_periodic(scheme=mol_b200.WENOScheme(), grid_x=examples.stretched_grid(0, 2, n + 1),
                                                          grid_y=examples.sinus_stretched_grid(0, 2, n + 1, 0.1)),
      "nonlin1d": lambda: examples.nonlinear_diffusion_1d(dx=1.0 / n)}[case]
t0 = time.perf_counter()
run = SlabRunner(*mk(), rank, world, local, weak=weak)
if os.environ.get("MOL_BENCH_GENERIC"):
    run.plan.set_option("kernel", 1)
t1 = time.perf_counter()
nbytes = run.state_len * 8
nbuf = max(2, int(3e8 // nbytes) + 1) if nbytes < 3e8 else 2
us = [torch.rand(run.state_len, dtype=torch.float64, device=dev) for _ in range(nbuf)]
dus = [torch.empty_like(us[0]) for _ in range(nbuf)]
for i in range(5):
    run.rhs(dus[i % nbuf], us[i % nbuf], 0.0)
torch.cuda.synchronize()
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
torch.cuda.synchronize()
l0 = run.launch_count()
e0.record()
for i in range(K):
    run.rhs(dus[i % nbuf], us[i % nbuf], 0.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
gbs = 2 * nbytes / (ms * 1e-3) / 1e9
if rank == 0:
    print(f"{case} n={n} world={world}: setup {t1-t0:.1f}s, {ms*1e3:.1f} us/RHS, {run.cells_local*world/(ms*1e-3):.3e} pts/s (all ranks), "
          f"{gbs:.0f} GB/s algorithmic per GPU = {gbs/6546.9:.3f} of measured copy rate, launches/RHS {(run.launch_count()-l0)/K:.1f}", flush=True)
if world > 1: dist.destroy_process_group()
