"""Slab decomposition across ranks: one process per GPU, the grid split along the
slowest-varying spatial axis, `radius` ghost planes exchanged with the two neighbours per RHS
evaluation (ring for a periodic split axis).  The reference has no multi-process path at all; this
is the B200-side extension the north star asks for.

torch.distributed is plumbing only (rendezvous + NCCL send/recv of the ghost planes); the stencil
kernels read the received planes directly through the plan's halo pointers.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .lowering import lower


class SlabRunner:
    """RHS evaluation of one rank's slab.  weak=True: every rank owns a full copy of the per-GPU
    problem size and the global problem is `world` slabs stacked along the split axis."""

    def __init__(self, pdesys, disc, rank=0, world=1, local_device=0, weak=True):
        import torch
        self.rank, self.world, self.device = rank, world, local_device
        self.torch = torch
        self.dev = torch.device("cuda", local_device)
        if world > 1 and weak:
            pdesys, disc = stack_domain(pdesys, disc, world)
        self.program = lower(pdesys, disc)
        self.plan = capi.Plan(self.program.text, local_device)
        self.nv = self.plan.nvar
        if world > 1:
            self._init_dist()
        else:
            self.state_len = self.plan.state_len
            self.cells_local = self.state_len // self.nv

    # -- world > 1 -----------------------------------------------------------------------------------
    def _init_dist(self):
        import torch.distributed as dist
        torch = self.torch
        self.dist = dist
        self.plan.dist_init(self.rank, self.world)
        info = self.plan.dist_info()
        self.H, self.plane, self.rows = info["radius"], info["plane"], info["rows_local"]
        self.state_len = self.nv * self.rows * self.plane
        self.cells_local = self.rows * self.plane
        self.halo_lo = torch.zeros(self.nv * self.H * self.plane, dtype=torch.float64, device=self.dev)
        self.halo_hi = torch.zeros_like(self.halo_lo)
        self.plan.dist_set_halo(self.halo_lo.data_ptr(), self.halo_hi.data_ptr())
        self.prev = (self.rank - 1) % self.world
        self.next = (self.rank + 1) % self.world
        self.periodic_split = info["periodic"]
        self.comm_stream = torch.cuda.Stream(self.dev)

    def exchange(self, u):
        """Ring exchange of the first/last H planes of every variable (NCCL send/recv, grouped)."""
        dist, torch = self.dist, self.torch
        U = u.view(self.nv, self.rows, self.plane)
        hl = self.halo_lo.view(self.nv, self.H, self.plane)
        hh = self.halo_hi.view(self.nv, self.H, self.plane)
        ops = []
        has_prev = self.periodic_split or self.rank > 0
        has_next = self.periodic_split or self.rank < self.world - 1
        for v in range(self.nv):
            if has_next:
                ops.append(dist.P2POp(dist.isend, U[v, self.rows - self.H:], self.next))
                ops.append(dist.P2POp(dist.irecv, hh[v], self.next))
            if has_prev:
                ops.append(dist.P2POp(dist.isend, U[v, :self.H], self.prev))
                ops.append(dist.P2POp(dist.irecv, hl[v], self.prev))
        return dist.batch_isend_irecv(ops) if ops else []

    def rhs(self, du, u, t):
        torch = self.torch
        if self.world == 1:
            self.plan.rhs(du.data_ptr(), u.data_ptr(), t, None, torch.cuda.current_stream(self.dev).cuda_stream)
            return
        cur = torch.cuda.current_stream(self.dev)
        self.comm_stream.wait_stream(cur)                # u must be complete before it is sent
        with torch.cuda.stream(self.comm_stream):
            reqs = self.exchange(u)
        # interior tiles need no ghost planes: they run while the planes are in flight
        self.plan.rhs_part(du.data_ptr(), u.data_ptr(), t, capi.PART_INTERIOR, cur.cuda_stream)
        for r in reqs:
            r.wait()
        cur.wait_stream(self.comm_stream)
        self.plan.rhs_part(du.data_ptr(), u.data_ptr(), t, capi.PART_BOUNDARY, cur.cuda_stream)

    def launch_count(self):
        return self.plan.launch_count()

    def kernel_name(self):
        return "mol_rhs_tiled (TMA, double-buffered)" if self.program.corebox is not None else "mol_rhs_generic"

    def describe(self):
        if self.world == 1:
            return "single GPU"
        return (f"slab decomposition along the last axis over {self.world} ranks, {self.H} ghost plane(s)/side/variable "
                "by NCCL send/recv on a side stream, overlapped with interior tiles")


def stack_domain(pdesys, disc, world):
    """Weak scaling: stretch the last spatial axis `world` times (same spacing, `world` x the nodes)."""
    import copy
    from .interface import Interval
    sys2, disc2 = copy.copy(pdesys), copy.copy(disc)
    t = disc.time
    xs = [a for a in pdesys.dvs[0].args if a != t]
    last = xs[-1]
    doms = []
    lo = hi = None
    for iv in pdesys.domains:
        if iv.var == last:
            lo, hi = float(iv.lo), float(iv.hi)
            doms.append(Interval(iv.var, lo, lo + (hi - lo) * world))
        else:
            doms.append(iv)
    sys2.domains = doms
    # boundary conditions written at the old upper end move to the new one
    new_hi = lo + (hi - lo) * world

    def move(e):
        import sympy as sp
        reps = {}
        for call in e.atoms(sp.core.function.AppliedUndef):
            for dv in pdesys.dvs:
                if call.func == dv.func:
                    k = list(dv.args).index(last)
                    a = call.args[k]
                    if a.is_number and abs(float(a) - hi) < 1e-12:
                        args = list(call.args)
                        args[k] = sp.Float(new_hi)
                        reps[call] = call.func(*args)
        return e.xreplace(reps)
    from .interface import Equation
    sys2.bcs = [Equation(move(b.lhs), move(b.rhs)) for b in pdesys.bcs]
    spec = disc.dxs[last]
    disc2.dxs = dict(disc.dxs)
    if isinstance(spec, (int, np.integer)):
        disc2.dxs[last] = (int(spec) - 1) * world + 1
    elif np.ndim(spec) > 0:
        raise NotImplementedError("weak scaling of explicit non-uniform grids")
    return sys2, disc2
