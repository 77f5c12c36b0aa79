#!/usr/bin/env python
"""bench.py — Brusselator-2D RHS grid-point updates/s and % of HBM roofline (BASELINE.json metric).

A "step" is one evaluation of the semi-discrete RHS f!(du,u,p,t) over the 4096^2 two-species
periodic Brusselator (BASELINE.json configs[1], the configuration the metric is quoted on).
1 grid-point update = one (i,j) cell, both species = 32 algorithmic bytes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 4096] [--impl reference]

`value`   : device-resident throughput (inputs already in HBM), CUDA events on the launch stream.
`e2e`     : same metric through the host-buffer call (pinned host u -> H2D -> RHS -> D2H du).
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
    RHS (oracle/bruss_ref.c) with every host thread, K steps of one RHS evaluation each."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    N = args.size
    nthreads = cref.lib().bruss_ref_max_threads()
    rng = np.random.default_rng(0)
    u = rng.uniform(0.0, 3.0, 2 * N * N)
    du = np.empty_like(u)
    g = np.arange(N + 1) / N
    steps = min(args.steps, 50)
    for _ in range(min(args.warmup, 3)):
        cref.bruss_rhs(u, g, g, N, 0.0, nthreads=nthreads, out=du)
    t0 = time.perf_counter()
    for _ in range(steps):
        cref.bruss_rhs(u, g, g, N, 0.0, nthreads=nthreads, out=du)
    el = time.perf_counter() - t0
    val = N * N * steps / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"brusselator2d_{N}x{N}_periodic_2species_rhs", "size": N},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                             "sample": f"{steps} RHS evaluations at {N}^2 (C restatement of the reference's generated RHS, OpenMP)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import _mol_import  # noqa: F401
    import mol_b200
    from mol_b200 import capi, examples

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
    us = [torch.from_numpy(rng.uniform(0.0, 3.0, n_loc)).to(dev) for _ in range(nbuf)]
    dus = [torch.empty_like(us[0]) for _ in range(nbuf)]
    stream = torch.cuda.current_stream(dev)

    def step(i):
        runner.rhs(dus[i % nbuf], us[i % nbuf], 0.0)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    l0 = runner.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with (ClockSampler(local) if rank == 0 else contextlib.nullcontext()) as clk:
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(K):
            step(i)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = runner.launch_count() - l0
    if dist:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        dist.barrier()
    updates_per_rank = runner.cells_local
    value = updates_per_rank * world * K / (ms * 1e-3)

    # ---- e2e: host buffers through the reference-facing call (H2D + RHS + D2H inside the timed region)
    hu = torch.from_numpy(rng.uniform(0.0, 3.0, n_loc)).pin_memory()
    hdu = torch.empty(n_loc, dtype=torch.float64).pin_memory()
    Ke = max(3, min(K, 10))

    def e2e_step():
        if world == 1:      # the library's host-buffer call: chunked H2D / sweep / D2H pipeline (mol_rhs_host)
            runner.plan.rhs_host(hdu.data_ptr(), hu.data_ptr(), 0.0, None, 0, stream.cuda_stream)
        else:               # slab mode keeps the state resident; host buffers go through explicit copies
            us[0].copy_(hu, non_blocking=True)
            runner.rhs(dus[0], us[0], 0.0)
            hdu.copy_(dus[0], non_blocking=True)

    e2e_step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0.record(stream)
    for _ in range(Ke):
        e2e_step()
    e1.record(stream)
    torch.cuda.synchronize()
    ems = e0.elapsed_time(e1)
    if dist:
        tms = torch.tensor([ems], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ems = float(tms.item())
    e2e_val = updates_per_rank * world * Ke / (ems * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_ms = ms / K
        achieved = updates_per_rank * BYTES_PER_UPDATE / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"brusselator2d_{N}x{N}_periodic_2species_rhs", "size": N,
                       "per_gpu_cells": updates_per_rank, "parallelism": runner.describe(),
                       "l2_policy": f"inputs larger than L2: {nbuf} rotating (u,du) sets of {2 * n_loc * 8 / 1e6:.0f} MB each",
                       "kernel": runner.kernel_name()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": updates_per_rank * BYTES_PER_UPDATE,
                         "kernel_ms": per_launch_ms},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n_loc * 8, "d2h_bytes_per_step": n_loc * 8,
                    "steps": Ke},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        traffic_file = os.path.join(ROOT, "profiles", "r01_tiled_dram_bytes.json")
        if os.path.exists(traffic_file):
            try:
                line["roofline"]["traffic"] = json.load(open(traffic_file)).get(f"N{N}")
            except Exception:
                pass
        if world == 1:
            from oracle import cref
            nth = cref.lib().bruss_ref_max_threads()
            v1, n1, el1 = cpu_restatement(N, args.cpu_seconds / 2, 1)
            vn, nn, eln = cpu_restatement(N, args.cpu_seconds / 2, nth)
            line["cpu_baseline"] = {"value": vn, "unit": UNIT, "cores": nth, "kind": "port",
                                    "value_1thread": v1,
                                    "sample": f"{nn} RHS evaluations at {N}^2 in {eln:.1f}s on {nth} threads (+{n1} on 1 thread in {el1:.1f}s); "
                                              "C restatement of the reference's generated RHS (oracle/bruss_ref.c)"}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
