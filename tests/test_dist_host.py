"""Host-side logic of the slab decomposition on CPU: partition arithmetic (C ABI, no GPU), slab
geometry from a compile-only plan, and the ghost-plane exchange protocol over gloo with world_size 2
and 3 (ring + open ends, including the two-rank ring where prev == next)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)            # spawned workers import this module without conftest.py
import _mol_import  # noqa: E402,F401
import mol_b200  # noqa: E402
from mol_b200 import capi
import problems as examples
from mol_b200.distributed import exchange_planes, stack_domain  # noqa: E402


def test_partition_covers_everything_once():
    for n in (7, 64, 4096, 1023):
        for p in (1, 2, 3, 4, 8):
            got = [capi.dist_partition(n, p, r) for r in range(p)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (a, c), (b, _) in zip(got, got[1:]):
                assert a + c == b
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_slab_geometry_of_compile_only_plans():
    sys_, disc = stack_domain(*examples.brusselator_2d(64), 2)
    prog = mol_b200.symbolic_discretize(sys_, disc)
    assert prog.nstate == 2 * 64 * 128
    for rank in (0, 1):
        plan = capi.Plan(prog.text, device=-1)
        plan.dist_init(rank, 2)
        info = plan.dist_info()
        assert (info.halo_planes, info.plane_len, info.n_planes, info.first_plane) == (1, 64, 64, 64 * rank)
        assert info.periodic == 1 and info.prev_rank == info.next_rank == 1 - rank
        assert plan.state_len == info.state_len_local == 2 * 64 * 64 and info.halo_len == 2 * 1 * 64
        assert len(plan.cubin("tiled_nin1_tma_dist")) > 0          # slab variants compile for sm_100a without a GPU
        plan.close()
    sys_, disc = examples.diffusion_reaction_3d(n=32, periodic=False)
    plan = capi.Plan(mol_b200.symbolic_discretize(sys_, disc).text, device=-1)
    plan.dist_init(0, 2)
    info = plan.dist_info()
    assert info.periodic == 0 and info.prev_rank == -1 and info.next_rank == 1 and info.plane_len == 32 * 32
    plan.close()
    # 1-D problems do not shard
    plan = capi.Plan(mol_b200.symbolic_discretize(*examples.heat_1d_dirichlet(dx=0.01)).text, device=-1)
    with pytest.raises(capi.MolError):
        plan.dist_init(0, 2)
    plan.close()


def _worker(rank, world, port, periodic, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nvar, rows, plane, H = 2, 6, 5, 2
    # global field value = 1000*var + global plane index + 0.01*in-plane index
    first = rank * rows
    U = torch.empty(nvar, rows, plane, dtype=torch.float64)
    for v in range(nvar):
        for r in range(rows):
            U[v, r] = 1000 * v + (first + r) + 0.01 * torch.arange(plane, dtype=torch.float64)
    hl = torch.full((nvar * H * plane,), -1.0, dtype=torch.float64)
    hh = torch.full((nvar * H * plane,), -1.0, dtype=torch.float64)
    prev = rank - 1 if rank > 0 else (world - 1 if periodic else -1)
    nxt = rank + 1 if rank < world - 1 else (0 if periodic else -1)
    for req in exchange_planes(dist, U.view(-1), hl, hh, nvar, rows, plane, H, prev, nxt):
        req.wait()
    total = world * rows
    ok = True
    for v in range(nvar):
        for k in range(H):
            lo_plane = (first - H + k) % total
            hi_plane = (first + rows + k) % total
            exp_lo = 1000 * v + lo_plane + 0.01 * torch.arange(plane, dtype=torch.float64)
            exp_hi = 1000 * v + hi_plane + 0.01 * torch.arange(plane, dtype=torch.float64)
            got_lo = hl.view(nvar, H, plane)[v, k]
            got_hi = hh.view(nvar, H, plane)[v, k]
            ok &= bool(torch.equal(got_lo, exp_lo)) if prev >= 0 else bool((got_lo == -1).all())
            ok &= bool(torch.equal(got_hi, exp_hi)) if nxt >= 0 else bool((got_hi == -1).all())
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic", [(2, True), (2, False), (3, True)])
def test_ghost_plane_exchange_over_gloo(world, periodic):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + world * 2 + int(periodic)
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
