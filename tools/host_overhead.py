"""Diagnostic: host-side enqueue time per RHS call vs device time (run under torchrun for N > 1)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _mol_import  # noqa
import torch
import problems as examples
from mol_b200.distributed import SlabRunner

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
run = SlabRunner(*examples.brusselator_2d(N), rank, world, local, weak=True)
us = [torch.rand(run.state_len, dtype=torch.float64, device=dev) * 3 for _ in range(3)]
dus = [torch.empty_like(us[0]) for _ in range(3)]
for i in range(10):
    run.rhs(dus[i % 3], us[i % 3], 0.0)
torch.cuda.synchronize()
K = 300
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter(); e0.record()
for i in range(K):
    run.rhs(dus[i % 3], us[i % 3], 0.0)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"rank {rank}: host enqueue {1e6*(t1-t0)/K:.1f} us/call, device {1e3*e0.elapsed_time(e1)/K:.1f} us/call, wall {1e6*(t2-t0)/K:.1f} us/call", flush=True)
if world > 1: dist.destroy_process_group()
