"""Slab decomposition across ranks: one process per GPU, the grid split along the
slowest-varying spatial axis, `halo_planes` ghost planes exchanged with the two neighbours per RHS
evaluation (ring for a periodic split axis).  The reference has no multi-process path at all; this
is the B200-side extension the north star asks for.

The exchange itself lives in libmol_cuda.so (transport "nccl": ncclSend/ncclRecv on a private
stream, overlapped with the interior part of the sweep — see csrc/mol_dist.cpp).  torch.distributed
is plumbing only: rendezvous and the broadcast of the 128-byte NCCL unique id.  Transport "torch"
moves the planes with torch.distributed P2P ops instead (any backend; used by the CPU/gloo tests of
the exchange logic and available as a diagnostic on GPUs).
"""
from __future__ import annotations

import numpy as np

from . import capi
from .lowering import lower


def exchange_planes(dist, state, halo_lo, halo_hi, nvar, n_planes, plane_len, H, prev, next_):
    """Ghost-plane exchange on torch tensors (CPU or CUDA): the first H planes of every variable go to
    `prev`'s upper ghosts, the last H planes to `next_`'s lower ghosts (None/-1 = domain edge).
    Posting order: sends "top -> next", "bottom -> prev", then receives "lower <- prev",
    "upper <- next", so that with two ranks on a ring (prev == next) the k-th send pairs with the
    peer's k-th receive."""
    U = state.view(nvar, n_planes, plane_len)
    hl = halo_lo.view(nvar, H, plane_len)
    hh = halo_hi.view(nvar, H, plane_len)
    has_prev = prev is not None and prev >= 0
    has_next = next_ is not None and next_ >= 0
    ops = []
    for v in range(nvar):
        if has_next:
            ops.append(dist.P2POp(dist.isend, U[v, n_planes - H:].contiguous(), next_))
    for v in range(nvar):
        if has_prev:
            ops.append(dist.P2POp(dist.isend, U[v, :H].contiguous(), prev))
    for v in range(nvar):
        if has_prev:
            ops.append(dist.P2POp(dist.irecv, hl[v], prev))
    for v in range(nvar):
        if has_next:
            ops.append(dist.P2POp(dist.irecv, hh[v], next_))
    return dist.batch_isend_irecv(ops) if ops else []


class SlabRunner:
    """RHS evaluation of one rank's slab.  weak=True: every rank owns a full copy of the per-GPU
    problem size and the global problem is `world` slabs stacked along the split axis."""

    def __init__(self, pdesys, disc, rank=0, world=1, local_device=0, weak=True, transport="nccl"):
        import torch
        self.rank, self.world, self.device = rank, world, local_device
        self.torch = torch
        self.dev = torch.device("cuda", local_device)
        self.transport = transport
        if world > 1 and weak:
            pdesys, disc = stack_domain(pdesys, disc, world)
        self.program = lower(pdesys, disc)
        if self.program.segments is not None:
            raise NotImplementedError("slab decomposition splits one shared grid: systems whose variables live on different "
                                      "(interface-joined) domains run on one GPU")
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
        self.info = info
        self.H, self.plane, self.rows = info.halo_planes, info.plane_len, info.n_planes
        self.state_len = info.state_len_local
        self.cells_local = self.rows * self.plane
        self.prev, self.next = info.prev_rank, info.next_rank
        if self.transport == "nccl":
            # rank 0 creates the NCCL unique id; torch.distributed carries the 128 bytes
            box = [capi.dist_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.plan.dist_comm_init(box[0])
        else:
            self.halo_lo = torch.zeros(info.halo_len, dtype=torch.float64, device=self.dev)
            self.halo_hi = torch.zeros_like(self.halo_lo)
            self.plan.dist_set_halo(self.halo_lo.data_ptr(), self.halo_hi.data_ptr())
            self.comm_stream = torch.cuda.Stream(self.dev)

    def local_slice(self, global_state):
        """This rank's part of a global state vector (host numpy, variable-major)."""
        if self.world == 1:
            return global_state
        n_glob = self.info.state_len_global // self.nv
        G = np.asarray(global_state).reshape(self.nv, n_glob // self.plane, self.plane)
        a = self.info.first_plane
        return np.ascontiguousarray(G[:, a:a + self.rows]).reshape(-1)

    def rhs(self, du, u, t):
        torch = self.torch
        cur = torch.cuda.current_stream(self.dev)
        if self.world == 1 or self.transport == "nccl":
            self.plan.rhs(du.data_ptr(), u.data_ptr(), t, None, cur.cuda_stream)
            return
        self.comm_stream.wait_stream(cur)                # u must be complete before it is sent
        with torch.cuda.stream(self.comm_stream):
            reqs = exchange_planes(self.dist, u, self.halo_lo, self.halo_hi, self.nv, self.rows, self.plane,
                                   self.H, self.prev, self.next)
        # interior tiles need no ghost planes: they run while the planes are in flight
        self.plan.rhs_part(du.data_ptr(), u.data_ptr(), t, capi.PART_INTERIOR, None, cur.cuda_stream)
        for r in reqs:
            r.wait()
        cur.wait_stream(self.comm_stream)
        self.plan.rhs_part(du.data_ptr(), u.data_ptr(), t, capi.PART_BOUNDARY, None, cur.cuda_stream)

    def launch_count(self):
        return self.plan.launch_count()

    def kernel_name(self):
        return "mol_rhs_tiled (TMA, multi-stage)" if self.program.corebox is not None else "mol_rhs_generic"

    def describe(self):
        if self.world == 1:
            return "single GPU"
        how = self.plan.dist_transport() if self.transport == "nccl" else "torch.distributed P2P ops"
        return (f"slab decomposition along the last axis over {self.world} ranks, {self.H} ghost plane(s)/side/variable; "
                f"transport: {how}, on a side stream, overlapped with the interior tiles")


def stack_domain(pdesys, disc, world):
    """Weak scaling: stretch the last spatial axis `world` times (same spacing, `world` x the nodes)."""
    import copy
    from .interface import Equation, Interval
    import sympy as sp
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
    new_hi = lo + (hi - lo) * world

    def move(e):
        """boundary conditions written at the old upper end move to the new one"""
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
    sys2.bcs = [Equation(move(b.lhs), move(b.rhs)) for b in pdesys.bcs]
    spec = disc.dxs[last]
    disc2.dxs = dict(disc.dxs)
    if isinstance(spec, (int, np.integer)):
        disc2.dxs[last] = (int(spec) - 1) * world + 1
    elif np.ndim(spec) > 0:
        raise NotImplementedError("weak scaling of explicit non-uniform grids")
    return sys2, disc2
