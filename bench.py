#!/usr/bin/env python
"""bench.py — Brusselator-2D RHS grid-point updates/s and % of HBM roofline (BASELINE.json metric).

A "step" is one evaluation of the semi-discrete RHS f!(du,u,p,t) over the 4096^2 two-species
periodic Brusselator (BASELINE.json configs[1], the configuration the metric is quoted on).
1 grid-point update = one (i,j) cell, both species = 32 algorithmic bytes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 4096] [--impl reference]

`value`   : device-resident throughput (inputs already in HBM), CUDA events on the launch stream.
`e2e`     : same metric through the user-facing solve call with host buffers (pinned u0 -> H2D -> 10 Tsit5 steps
            -> D2H u(t1)); `e2e_rhs_host` = one f!(du,u,p,t) with host arrays per step (PCIe-bound).
`roofline`: algorithmic bytes / measured kernel time vs MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the oracle's C restatement of the reference's generated RHS on the host cores.
`--impl reference`: the same CPU restatement timed as its own arm (the reference is pure Julia and
cannot run in this image: kind = "port").
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "brusselator2d_rhs_gridpoint_updates_per_s"
UNIT = "grid-point updates/s"
BYTES_PER_UPDATE = 32.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML every 10 ms; nvidia-smi as a fallback).
    Only rank 0 samples: eight ranks polling NVML at a high rate compete with the launching threads for the host cores
    (and for the driver's lock) and show up as launch jitter in a 90 us step."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.sm, self.reasons, self.sm_max, self.stop, self.index = [], set(), None, False, index
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        for bit, name in ((n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                          (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6 and r[0].isdigit():
            self.sm.append(int(r[0]))
            self.sm_max = int(r[1]) if r[1].isdigit() else self.sm_max
            for k, name in enumerate(self.NAMES):
                if r[2 + k].lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self.stop:
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            time.sleep(0.01 if self.nvml else 0.05)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def workload_config(N, world):
    """`config` is a function of the workload only (size, ranks), so both arms print the same dict."""
    return {"workload": f"brusselator2d_{N}x{N}_periodic_2species_rhs", "size": N,
            "per_gpu_cells": N * N, "global_cells": N * N * world,
            "t": 0.0, "ic": "numpy default_rng(rank).uniform(0,3)",
            "l2_policy": f"inputs larger than L2: 3 rotating (u,du) sets of {2 * 2 * N * N * 8 / 1e6:.0f} MB each"}


def cpu_restatement(N, seconds, nthreads):
    """Times oracle/bruss_ref.c (CPU baseline leg: the only place bench.py executes oracle/)."""
    from oracle import cref
    rng = np.random.default_rng(0)
    u = rng.uniform(0.0, 3.0, 2 * N * N)
    du = np.empty_like(u)
    g = np.arange(N + 1) / N
    cref.bruss_rhs(u, g, g, N, 0.0, nthreads=nthreads, out=du)          # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        cref.bruss_rhs(u, g, g, N, 0.0, nthreads=nthreads, out=du)
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 10000:
            break
    return N * N * n / el, n, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  MethodOfLines.jl is pure
    Julia (no Julia in this image), so this arm times the C restatement of its generated
    RHS (oracle/bruss_ref.c) on every host core this process may use (sched_getaffinity, NOT
    OMP_NUM_THREADS: torchrun sets that to 1), W warm-up + K timed steps of one RHS evaluation each.
    Under torchrun only rank 0 works.  At N > 1 the workload is N stacked 4096^2 slabs; a step is a
    bounded sample of it (one slab's worth of cells; the RHS cost per cell does not depend on the slab)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    N = args.size
    nthreads = cref.host_threads()
    rng = np.random.default_rng(0)
    u = rng.uniform(0.0, 3.0, 2 * N * N)
    du = np.empty_like(u)
    g = np.arange(N + 1) / N
    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        cref.bruss_rhs(u, g, g, N, 0.0, nthreads=nthreads, out=du)
    t0 = time.perf_counter()
    for _ in range(K):
        cref.bruss_rhs(u, g, g, N, 0.0, nthreads=nthreads, out=du)
    el = time.perf_counter() - t0
    val = N * N * K / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * el / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(N, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                             "sample": f"{K} RHS evaluations of one {N}^2 slab (C restatement of the reference's generated RHS, "
                                       f"OpenMP over {nthreads} threads = every core in this process's affinity mask)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic"])
    ap.add_argument("--no-extra", action="store_true", help="skip the config-5 / Tsit5 records in `extra`")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import _mol_import  # noqa: F401
    import mol_b200
    from mol_b200 import capi
    import problems as examples

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    N = args.size

    from mol_b200 import distributed as mdist
    sys_, disc = examples.brusselator_2d(N)
    runner = mdist.SlabRunner(sys_, disc, rank, world, local, weak=True)
    n_loc = runner.state_len
    rng = np.random.default_rng(rank)
    nbuf = 3                                            # rotate buffer sets: 3 x (u, du) >> 126 MB L2
    hus = [rng.uniform(0.0, 3.0, n_loc) for _ in range(nbuf)]
    us = [torch.from_numpy(h).to(dev) for h in hus]
    dus = [torch.empty_like(us[0]) for _ in range(nbuf)]
    stream = torch.cuda.current_stream(dev)
    align = torch.zeros(1, dtype=torch.float64, device=dev)

    def max_over_ranks(x):
        if not dist:
            return float(x)
        tms = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        return float(tms.item())

    def timed(step, K, W):
        """W warm-up steps, then exactly K steps between two events on the launch stream, bracketed by
        synchronize + barrier + synchronize on both sides.  Every step also records its own event (per-step
        spread).  With several ranks a tiny all-reduce through the library's communicator is queued on the
        launch stream right before the first event: the ranks' device timelines then start together, whatever
        the host-side skew after the barrier."""
        for i in range(W):
            step(i)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        if dist:
            runner.plan.dist_allreduce_sum(align.data_ptr(), 1, stream.cuda_stream)
        evs[0].record(stream)
        for i in range(K):
            step(W + i)
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        ms = evs[0].elapsed_time(evs[K])
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
        return ms, per

    # rank 0 samples the clocks; the sampler (NVML init, thread start) is up BEFORE the barrier that opens the timed region
    clk = ClockSampler(local) if rank == 0 else None
    if clk:
        clk.__enter__()
    l0 = [0]

    def step(i):
        if i == W:
            l0[0] = runner.launch_count()
        runner.rhs(dus[i % nbuf], us[i % nbuf], 0.0)

    ms_local, per = timed(step, K, W)
    launches = runner.launch_count() - l0[0]
    if clk:
        clk.__exit__()
    ms = max_over_ranks(ms_local)
    med, p95, first = pct(per, 0.5), pct(per, 0.95), per[0]
    med_max = max_over_ranks(med)
    updates_per_rank = runner.cells_local
    value = updates_per_rank * world * K / (ms * 1e-3)

    # ---- forcing active (t = 2.0): same sweep, the indicator term switched on
    K2 = min(K, 50)
    ms_t2, _ = timed(lambda i: runner.rhs(dus[i % nbuf], us[i % nbuf], 2.0), K2, 3)
    ms_t2 = max_over_ranks(ms_t2)

    # ---- parity of what was just timed: this rank's slab against the C restatement of the reference's RHS on
    # the stacked global problem (rows below / above the slab = the neighbouring ranks' edge rows)
    from oracle import cref
    rows = updates_per_rank // N
    first_row = runner.info.first_plane if world > 1 else 0
    xg = np.arange(N + 1) / N
    yrows = (np.arange(first_row, first_row + rows) + 1) / N
    mine = hus[0].reshape(2, rows, N)
    if world > 1:
        below = np.random.default_rng((rank - 1) % world).uniform(0.0, 3.0, n_loc).reshape(2, rows, N)
        above = np.random.default_rng((rank + 1) % world).uniform(0.0, 3.0, n_loc).reshape(2, rows, N)
    else:
        below = above = mine
    parity = 0.0
    for tt in (0.0, 2.0):
        runner.rhs(dus[0], us[0], tt)
        torch.cuda.synchronize()
        ref = cref.bruss_rhs_slab(hus[0], below[0, -1], above[0, 0], below[1, -1], above[1, 0], xg, yrows, N, rows, tt,
                                  nthreads=max(1, cref.host_threads() // world))
        got = dus[0].cpu().numpy()
        parity = max(parity, float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
    parity = max_over_ranks(parity)

    # ---- e2e: the call a user makes, with HOST buffers and the copies inside the timed region.
    # The user-facing call of this path is solve(prob, Tsit5()) (north_star: "evaluating the RHS ... and stepping it
    # with explicit Runge-Kutta", u resident in HBM in between), so one e2e step is: pinned host u0 -> H2D -> mol_rk_solve
    # over E2E_RK_STEPS fixed Tsit5 steps (6 RHS evaluations each, +1 to start the FSAL chain) -> D2H of u(t1) to pinned
    # host memory.  The metric stays RHS grid-point updates/s: cells x RHS evaluations / time.  Ten steps per transfer is
    # far fewer than any real solve takes (the explicit stability limit at 4096^2 is dt ~ 1e-9), i.e. conservative.
    # `e2e_rhs_host` beside it is the per-evaluation host round trip (f!(du, u, p, t) with host arrays): 32 B per update
    # over PCIe in both directions, which caps it near 3e9 updates/s whatever the kernel does.
    E2E_RK_STEPS, E2E_DT = 10, 1.0e-9
    hu = torch.from_numpy(hus[1]).pin_memory()
    hdu = torch.empty(n_loc, dtype=torch.float64).pin_memory()
    Ke = max(3, min(K, 5))
    rk = capi.RK(runner.plan, "tsit5", 1e-6, 1e-3)
    nf_box = [0]

    def e2e_step(i):
        us[0].copy_(hu, non_blocking=True)
        st = rk.solve(us[0].data_ptr(), 0.0, E2E_RK_STEPS * E2E_DT, E2E_DT, False, None, 0, 10 ** 6, stream.cuda_stream)
        nf_box[0] = int(st.nf)
        hdu.copy_(us[0], non_blocking=True)

    ems, _ = timed(e2e_step, Ke, 1)
    ems = max_over_ranks(ems)
    rk.close()
    e2e_val = updates_per_rank * world * nf_box[0] * Ke / (ems * 1e-3)

    def rhs_host_step(i):
        if world == 1:      # the library's host-buffer call: chunked H2D / sweep / D2H pipeline (mol_rhs_host)
            runner.plan.rhs_host(hdu.data_ptr(), hu.data_ptr(), 0.0, None, 0, stream.cuda_stream)
        else:               # slab mode keeps the state resident; host buffers go through explicit copies
            us[0].copy_(hu, non_blocking=True)
            runner.rhs(dus[0], us[0], 0.0)
            hdu.copy_(dus[0], non_blocking=True)

    hms, _ = timed(rhs_host_step, Ke, 1)
    hms = max_over_ranks(hms)
    rhs_host_val = updates_per_rank * world * Ke / (hms * 1e-3)

    extra = {}
    if not args.no_extra:
        extra = extra_records(torch, dist, runner, rank, world, local, dev, max_over_ranks)

    if parity > 1e-12:
        raise SystemExit(f"bench.py: RHS parity against the C restatement failed: max rel {parity:.3e} > 1e-12")

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_ms = ms / K
        achieved = updates_per_rank * BYTES_PER_UPDATE / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(N, world),
            "parallelism": runner.describe(), "kernel": runner.kernel_name(),
            "per_step_ms": {"median": med, "median_max_over_ranks": med_max, "p95": p95, "first": first,
                            "sum_over_K_median": ms_local / (K * med)},
            "rhs_t2_forcing_active": {"ms_per_step": ms_t2 / K2, "steps": K2,
                                      "frac": updates_per_rank * BYTES_PER_UPDATE / (ms_t2 / K2 * 1e-3) / 1e9 / peak},
            "dist_parity" if world > 1 else "parity": {
                "max_rel": parity, "ranks": world, "t": [0.0, 2.0], "bar": 1e-12,
                "against": "oracle/bruss_ref.c on the stacked global problem, every rank's slab, max over ranks"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": updates_per_rank * BYTES_PER_UPDATE,
                         "kernel_ms": per_launch_ms},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n_loc * 8, "d2h_bytes_per_step": n_loc * 8,
                    "steps": Ke, "ms_per_step": ems / Ke, "rhs_evals_per_step": nf_box[0],
                    "call": f"mol_rk_solve with host buffers: pinned u0 -> H2D -> {E2E_RK_STEPS} fixed Tsit5 steps "
                            f"(dt = {E2E_DT:g}) -> D2H u(t1); updates = cells x RHS evaluations"},
            "e2e_rhs_host": {"value": rhs_host_val, "unit": UNIT, "h2d_bytes_per_step": n_loc * 8,
                             "d2h_bytes_per_step": n_loc * 8, "steps": Ke,
                             "call": "mol_rhs_host: one f!(du, u, p, t) with host arrays per step (PCIe-bound: 32 B per update)"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        if extra:
            line["extra"] = extra
        traffic_file = os.path.join(ROOT, "profiles", "tiled_dram_bytes.json")
        if os.path.exists(traffic_file):
            try:
                tf = json.load(open(traffic_file))
                line["roofline"]["traffic"] = tf.get(f"N{N}")
                line["roofline"]["traffic_source"] = tf.get("source")
            except Exception:
                pass
        if world == 1:
            nth = cref.host_threads()
            v1, n1, el1 = cpu_restatement(N, args.cpu_seconds / 2, 1)
            vn, nn, eln = cpu_restatement(N, args.cpu_seconds / 2, nth)
            line["cpu_baseline"] = {"value": vn, "unit": UNIT, "cores": nth, "kind": "port",
                                    "value_1thread": v1,
                                    "sample": f"{nn} RHS evaluations at {N}^2 in {eln:.1f}s on {nth} threads (+{n1} on 1 thread in {el1:.1f}s); "
                                              "C restatement of the reference's generated RHS (oracle/bruss_ref.c)"}
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def extra_records(torch, dist, runner2d, rank, world, local, dev, max_over_ranks):
    """Second records the driver's runs capture beside the headline: BASELINE config 5 (3-D
    diffusion-reaction, 1024 x 1024 x 128 per GPU = 1024^3 over 8 GPUs) with slab parity, and one fused Tsit5 step of
    the 4096^2 Brusselator.  Failures are reported in the record, never raised: the headline must survive."""
    import _mol_import  # noqa: F401
    from mol_b200 import capi
    import problems as examples
    from mol_b200 import distributed as mdist
    from oracle import cref
    peak, _ = peaks()
    out = {}
    # -- config 5
    try:
        n, nzl = 1024, 128
        sys3, disc3 = examples.diffusion_reaction_3d(n=n, periodic=True, nz=nzl)
        run3 = mdist.SlabRunner(sys3, disc3, rank, world, local, weak=True)
        nl = run3.state_len
        rng = np.random.default_rng(100 + rank)
        h3 = rng.uniform(0.0, 1.0, nl)
        u3 = [torch.from_numpy(h3).to(dev), torch.rand(nl, dtype=torch.float64, device=dev)]
        d3 = [torch.empty_like(u3[0]) for _ in range(2)]
        for i in range(3):
            run3.rhs(d3[i % 2], u3[i % 2], 0.0)
        K3 = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(K3):
            run3.rhs(d3[i % 2], u3[i % 2], 0.0)
        e1.record()
        torch.cuda.synchronize()
        us3 = max_over_ranks(e0.elapsed_time(e1) / K3 * 1e3)
        # parity of this rank's slab (neighbouring slabs' edge planes regenerated from their seeds)
        run3.rhs(d3[0], u3[0], 0.0)
        torch.cuda.synchronize()
        P = n * n
        if world > 1:
            lo = np.random.default_rng(100 + (rank - 1) % world).uniform(0.0, 1.0, nl)[-P:]
            hi = np.random.default_rng(100 + (rank + 1) % world).uniform(0.0, 1.0, nl)[:P]
        else:
            lo, hi = h3[-P:], h3[:P]
        ref = cref.fisher3d_rhs_slab(h3, lo, hi, n, n, nzl, 1.0 / n, nthreads=max(1, cref.host_threads() // world))
        got = d3[0].cpu().numpy()
        par = max_over_ranks(float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
        out["fisher3d_1024x1024x128_per_gpu"] = {
            "us": us3, "frac": nl * 16.0 / (us3 * 1e-6) / 1e9 / peak, "bytes_per_point": 16, "steps": K3,
            "points_per_s_all_ranks": nl * world / (us3 * 1e-6), "parity_max_rel": par,
            "parity_against": "oracle/configs_ref.c fisher3d_ref_rhs_slab, every rank's slab", "ranks": world}
        del run3, u3, d3
        torch.cuda.empty_cache()
    except Exception as e:          # noqa: BLE001
        out["fisher3d_1024x1024x128_per_gpu"] = {"error": repr(e)[:300]}
    # -- slab Tsit5 (stage-combine-on-load with per-array ghost planes, all-reduced error norm) against the oracle's
    # Tsit5 on the GLOBAL problem driven by the C restatement: final state within the integrator tolerance
    if world > 1:
        try:
            from oracle.rk import solve_tsit5
            Nt, T, tol = 16 * world if world > 4 else 64, 2.0e-3, 1e-8
            syst, disct = examples.brusselator_2d(Nt, tmax=T)
            runt = mdist.SlabRunner(syst, disct, rank, world, local, weak=False)
            g = np.arange(Nt + 1) / Nt
            X, Y = np.meshgrid(g[1:], g[1:], indexing="xy")          # x fastest
            u0 = np.concatenate([(22.0 * (Y * (1 - Y)) ** 1.5).reshape(-1), (27.0 * (X * (1 - X)) ** 1.5).reshape(-1)])
            _, usol, stt = solve_tsit5(lambda uu, tt: cref.bruss_rhs(uu, g, g, Nt, tt), u0, (0.0, T), abstol=tol, reltol=tol)
            ul = torch.from_numpy(runt.local_slice(u0)).to(dev)
            rk = capi.RK(runt.plan, "tsit5", tol, tol)
            st = rk.solve(ul.data_ptr(), 0.0, T, 0.0, True, None, 0, 10 ** 6, torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.synchronize()
            ref = runt.local_slice(usol[-1])
            err = max_over_ranks(float(np.max(np.abs(ul.cpu().numpy() - ref) / (tol + tol * np.abs(ref)))))
            out["dist_tsit5_parity"] = {"size": Nt, "t_final": T, "abstol": tol, "reltol": tol, "ranks": world,
                                        "max_err_over_tol": err, "bar": 100.0, "retcode": int(st.retcode),
                                        "steps_gpu": int(st.naccept), "steps_oracle": int(stt["naccept"]),
                                        "ok": bool(st.retcode == 0 and err <= 100.0)}
            rk.close()
            del runt
        except Exception as e:      # noqa: BLE001
            out["dist_tsit5_parity"] = {"error": repr(e)[:300]}
    # -- fused Tsit5 step at 4096^2 (single-GPU record)
    if world == 1:
        try:
            plan = runner2d.plan
            n = plan.state_len
            u = torch.rand(n, dtype=torch.float64, device=dev) * 3
            st = torch.cuda.current_stream(dev).cuda_stream
            rk = capi.RK(plan, "tsit5", 1e-6, 1e-3)
            t, dt = 0.0, 1e-9
            for _ in range(3):
                t, _, _ = rk.step(u.data_ptr(), t, dt, adaptive=False, stream=st)
            torch.cuda.synchronize()
            Kr = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(Kr):
                t, _, _ = rk.step(u.data_ptr(), t, dt, adaptive=False, stream=st)
            e1.record()
            torch.cuda.synchronize()
            msr = e0.elapsed_time(e1) / Kr
            passes = 32          # 30 array passes of the fused step (DESIGN §4 RK) + the in-place form's u <- u+ copy
            out["tsit5_step_4096"] = {"ms": msr, "passes": passes, "rhs_per_step": 6,
                                      "frac": passes * n * 8 / (msr * 1e-3) / 1e9 / peak,
                                      "rhs_updates_per_s": 6 * (n // 2) / (msr * 1e-3)}
            rk.close()
        except Exception as e:      # noqa: BLE001
            out["tsit5_step_4096"] = {"error": repr(e)[:300]}
        # -- adaptive Tsit5 on a mid-size problem (512^2 Brusselator): step controller on the device, attempts queued as a
        # captured graph, against the host-driven loop (one read-back per attempt); same controller kernel in both
        try:
            import time as _time
            Nm = 512
            T = 400 * 0.3 * (1.0 / Nm) ** 2 / 10.0
            rec = {"size": Nm, "t_final": T}
            sysm, discm = examples.brusselator_2d(Nm, tmax=T)
            import mol_b200
            old = os.environ.get("MOL_RK_QUEUED")
            finals = {}
            for mode, key in (("1", "queued"), ("0", "host_loop")):
                os.environ["MOL_RK_QUEUED"] = mode
                probm = mol_b200.discretize(sysm, discm)
                um = torch.from_numpy(probm.u0).to(dev)
                rk = capi.RK(probm.plan, "tsit5", 1e-6, 1e-3)
                st = torch.cuda.current_stream(dev).cuda_stream
                best = None
                for _ in range(3):
                    um.copy_(torch.from_numpy(probm.u0))
                    torch.cuda.synchronize()
                    w0 = _time.perf_counter()
                    stt = rk.solve(um.data_ptr(), 0.0, T, 0.0, True, None, 0, 10 ** 6, st)
                    torch.cuda.synchronize()
                    el = _time.perf_counter() - w0
                    best = el if best is None else min(best, el)
                att = int(stt.naccept + stt.nreject)
                rec[key] = {"us_per_attempt": best / max(1, att) * 1e6, "naccept": int(stt.naccept), "nreject": int(stt.nreject),
                            "retcode": int(stt.retcode)}
                finals[key] = um.cpu().numpy()
                rk.close()
            if old is None:
                os.environ.pop("MOL_RK_QUEUED", None)
            else:
                os.environ["MOL_RK_QUEUED"] = old
            rec["max_abs_difference_of_final_states"] = float(np.max(np.abs(finals["queued"] - finals["host_loop"])))
            out["tsit5_adaptive_512"] = rec
        except Exception as e:      # noqa: BLE001
            out["tsit5_adaptive_512"] = {"error": repr(e)[:300]}
    return out


if __name__ == "__main__":
    main()
