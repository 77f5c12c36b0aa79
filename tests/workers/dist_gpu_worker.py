"""torchrun worker for tests/test_gpu_dist.py: slab-decomposed RHS / Tsit5 on WORLD_SIZE GPUs against
the oracle evaluated on the GLOBAL problem (rank-count invariance)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import _mol_import  # noqa: E402,F401
import mol_b200  # noqa: E402
from mol_b200 import capi
import problems as examples
from mol_b200.distributed import SlabRunner  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle.discretize import OracleProblem
    n3 = 24 if world <= 2 else 9 * world          # every rank needs >= 8 planes along the split axis
    cases = {
        "bruss_periodic": lambda: examples.brusselator_2d(48 * world),
        "burgers2d_bc": lambda: examples.burgers_2d(nx=40, ny=44),
        "fisher3d_periodic": lambda: examples.diffusion_reaction_3d(n=n3, periodic=True),
        "fisher3d_dirichlet_z": lambda: examples.diffusion_reaction_3d(n=n3, periodic=False),
        # several xy tiles of the z-marching kernel inside every slab
        "fisher3d_periodic_multitile": lambda: examples.diffusion_reaction_3d(n=72, periodic=True, nz=20 * world),
    }
    transports = ["nccl", "torch"]
    worst = 0.0
    for name, mk in cases.items():
        sys_, disc = mk()
        orc = OracleProblem(sys_, disc)
        rng = np.random.default_rng(11)
        ug = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
        for transport in transports:
            run = SlabRunner(sys_, disc, rank, world, local, weak=False, transport=transport)
            assert run.info.state_len_global == orc.nstate
            ul = torch.from_numpy(run.local_slice(ug)).to(dev)
            dul = torch.empty_like(ul)
            for t in (0.0, 0.37):
                ref = orc.rhs(ug, t)
                scale = float(np.max(orc.rhs_termscale(ug, t)))
                run.rhs(dul, ul, t)
                torch.cuda.synchronize()
                err = float(np.max(np.abs(dul.cpu().numpy() - run.local_slice(ref))))
                worst = max(worst, err / scale)
                assert err <= 1e-13 * scale, (name, transport, rank, t, err / scale)
                assert err <= 1e-12 * np.max(np.abs(ref)), (name, transport, rank, t)
    # distributed Tsit5 (stage-combine-on-load with per-array ghost planes, all-reduced error norm)
    sys_, disc = examples.brusselator_2d(32, tmax=0.05)
    orc = OracleProblem(sys_, disc)
    from oracle.rk import solve_tsit5
    ts, us, _ = solve_tsit5(orc.rhs, orc.u0, (0.0, 0.05), abstol=1e-9, reltol=1e-9)
    run = SlabRunner(sys_, disc, rank, world, local, weak=False)
    ul = torch.from_numpy(run.local_slice(orc.u0)).to(dev)
    rk = capi.RK(run.plan, "tsit5", 1e-9, 1e-9)
    st = rk.solve(ul.data_ptr(), 0.0, 0.05, 0.0, True, None, 0, 10 ** 5, torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    assert st.retcode == 0, st.retcode
    np.testing.assert_allclose(ul.cpu().numpy(), run.local_slice(us[-1]), rtol=1e-6, atol=1e-7)
    rk.close()
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world} worst_err_over_termscale={worst:.3e} tsit5_steps={st.naccept}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
