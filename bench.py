#!/usr/bin/env python
"""bench.py — Mcell-updates/s per mom_step! on the 3-D Taylor-Green vortex (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl reference]

One JSON line on stdout (rank 0).  `value` is whole-job throughput with all fields resident in
HBM (CUDA events on the handle's stream around exactly K wl_mom_step calls); `e2e` is the same
metric through the public API (`Simulation` fields uploaded from pinned host memory, K ×
`sim_step`, velocity and pressure downloaded) with the copies inside the timed region;
`roofline` is the dominant kernel class timed live with CUDA events inside the library;
`cpu_baseline` is the CPU oracle (a port: the reference is Julia and cannot run here) on a bounded
sample.  `--impl reference` times that oracle alone, with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mcell-updates/s per mom_step! (3D TGV periodic, Float32)"
RE = 1600.0


def tgv_ulam(N):
    kappa = 2 * np.pi / N

    def ulam(i, x):
        X, Y, Z = x[..., 0] * kappa, x[..., 1] * kappa, x[..., 2] * kappa
        if i == 1:
            return -np.sin(X) * np.cos(Y) * np.cos(Z)
        if i == 2:
            return np.cos(X) * np.sin(Y) * np.cos(Z)
        return np.zeros((1, 1, 1), dtype=np.float32)
    return ulam, 1.0 / (kappa * RE)


def workload_name(N):
    """`config.workload` — the same string on both arms (the driver compares them)."""
    return f"3D Taylor-Green vortex {N}^3 periodic, Re=1600, Float32, MultiLevelPoisson"


def bench_config(N, gpus):
    """`config` — identical on both arms for the same command line (the driver compares them); run-dependent details
    (V-cycle counts, level count, collective count, CPU sample) go into the top-level `details` key."""
    par = "1 GPU" if gpus == 1 else (f"z-slab x{gpus}: {N // gpus} planes/rank, peer-memory (NVLink) halo exchange and all-reduce from inside "
                                     "the kernels, NCCL for bootstrap only, coarse levels replicated")
    return {"workload": workload_name(N), "l2": "working set (~17 GB at 512^3) far exceeds the 126 MB L2; no flush needed", "parallelism": par}


def make_tgv(cls, N, **kw):
    ulam, nu = tgv_ulam(N)
    return cls((N, N, N), (0, 0, 0), N, U=1, nu=nu, ulam=ulam, perdir=(1, 2, 3), **kw)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_oracle_run(N, steps, warmup):
    """mom_step! of the CPU oracle on TGV N³, all host threads; returns Mcell-updates/s."""
    from oracle import oracle
    from oracle.oracle import OracleSimulation
    from waterlily_b200 import sim_step
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    oracle.set_num_threads(ncpu)   # all host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun: 1)
    sim = make_tgv(OracleSimulation, N)
    for _ in range(warmup):
        sim_step(sim, remeasure=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim_step(sim, remeasure=False)
    dt = time.perf_counter() - t0
    return N ** 3 * steps / dt / 1e6, dt / steps * 1e3, oracle.num_threads(), list(sim.pois.n[-2 * steps:])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.cpu_size
    val, ms, cores, ns = cpu_oracle_run(N, args.steps, args.warmup)
    sample = f"TGV {N}^3 periodic Re=1600 (bounded sample of the {args.size}^3 workload), {args.steps} mom_step! after {args.warmup} warm-up, V-cycles {ns}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.size, args.gpus), "details": {"sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mcell-updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "the reference is Julia and cannot run in this image; this is the C++/OpenMP oracle, a literal un-fused restatement of it"},
        "e2e": {"value": val, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    from waterlily_b200 import Simulation, sim_step, _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    if args.gpus > 1 and world == 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    N, K, W = args.size, args.steps, args.warmup
    dd = None
    if world > 1:
        import torch.distributed as dist
        from waterlily_b200 import lib
        dist.init_process_group("gloo")   # host-side plumbing only: id broadcast, barriers, max over ranks
        ident = [_cabi.nccl_unique_id(lib.load()) if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        dd = (rank, world, ident[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sim = make_tgv(Simulation, N, device=local, dist=dd)
    eng = sim._eng
    info = eng.dist_info()
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    # pinned host copies of the initial state for the e2e arm
    u_host = torch.from_numpy(np.ascontiguousarray(sim.flow.u.numpy().reshape(-1, order="F"))).pin_memory()
    p_host = torch.zeros(int(np.prod(eng.Ng)), dtype=torch.float32).pin_memory()
    U = (0.0, 0.0, 0.0)

    # ---- device-resident arm: W warm-up, then exactly K steps between CUDA events --------------
    eng.profile(True)          # the warm-up steps also create the library's CUDA-event pool
    for _ in range(W):
        eng.mom_step(U)
    eng.profile(True)          # reset the per-kernel statistics; events are reused
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count()
    c0 = eng.dist_info()["collectives"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    ns = []
    for _ in range(K):
        _, n1, n2 = eng.mom_step(U)
        ns += [n1, n2]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = sum_over_ranks(eng.launch_count() - l0)
    collectives = eng.dist_info()["collectives"] - c0
    value = N ** 3 * K / (ms * 1e-3) / 1e6
    names = eng.profile_classes()
    prof = {}
    for i, name in enumerate(names):
        pms, pl, pb = eng.profile_read(i)
        prof[name] = {"ms": pms, "launches": pl, "bytes": pb}
    eng.profile(False)
    peak, peak_src = measured_peak_hbm()
    kern = {k: v for k, v in prof.items() if k not in ("solver", "mom_step", "project_misc", "vcycle_tail") and v["launches"] > 0 and v["ms"] > 0}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    d = kern[dom]
    achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture at 512^3 on one GPU
    # (profiles/r02_ncu_full_top_kernels_512_c.txt, profiles/r02_launches_tgv512_1gpu_c.txt).  A class with two kernels
    # (conv_diff!: predictor / corrector launch) carries their mean, like `achieved`.
    ncu_traffic = {"conv_diff+bdim1": 0.5 * ((1.839e9 + 1.574e9) + (3.518e9 + 1.606e9)), "pcg_update": 2.216e9 + 1.050e9, "pcg_mult": 0.607e9 + 0.507e9,
                   "residual": 1.172e9 + 1.200e9, "increment": 1.262e9 + 1.577e9}
    traffic = ncu_traffic.get(dom) if (N == 512 and world == 1) else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "note": "dominant kernel = the profiling class with the largest device time in the timed region; conv_diff! (k_conv_diff3_tma2, BDIM! folded "
                        "into its store: 6 / 9 algorithmic words per cell) is bound by instruction issue, not HBM (1.1e9 warp instructions per launch at 64 % "
                        "issue-slot utilisation, ALU pipe 51 %, profiles/r02_ncu_full_top_kernels_512_c.txt); the streaming solver kernels run at "
                        "84-96 % of the measured HBM peak (kernels{})", "launches": d["launches"], "avg_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / prof["mom_step"]["ms"] if prof["mom_step"]["ms"] else None,
                "kernels": {k: {"GB/s": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else None, "ms_total": v["ms"], "launches": v["launches"]} for k, v in prof.items()}}

    # ---- e2e arm: host buffers in, K sim_step through the public API, host buffers out ---------
    Ng = eng.Ng
    u_np = u_host.numpy().reshape(Ng + (3,), order="F")
    p_np = p_host.numpy().reshape(Ng, order="F")
    out_u = torch.empty(u_host.shape, dtype=torch.float32, pin_memory=True)   # empty_like would drop the pinning
    out_p = torch.empty(p_host.shape, dtype=torch.float32, pin_memory=True)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.flow.u.set(u_np)               # the state a user hands over: velocity and pressure (u⁰ is not an input: the reference's Flow
    sim.flow.p.set(p_np)               # constructor makes it on the device with copy(u), and mom_step! opens with u⁰ .= u, Flow.jl:155)
    sim.flow.dt[:] = [np.float32(0.25)]
    eng.set_dt(0.25)
    for _ in range(K):
        sim_step(sim)                      # default remeasure=True: measure! (NoBody → no-op) + update!(pois) + mom_step!
    frac = (info["nz_local"] + 2) / Ng[2]   # a slab handle copies only its own planes of the global host array
    h2d = sum_over_ranks((u_host.numel() + p_host.numel()) * 4 * frac)
    # results back into pinned host memory
    import ctypes
    eng.cabi.check(eng.cabi.download(eng.h, _cabi.U, ctypes.cast(out_u.data_ptr(), _cabi.c_flt_p), out_u.numel()))
    eng.cabi.check(eng.cabi.download(eng.h, _cabi.P, ctypes.cast(out_p.data_ptr(), _cabi.c_flt_p), out_p.numel()))
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
    d2h = sum_over_ranks((out_u.numel() + out_p.numel()) * 4 * frac)
    e2e = {"value": N ** 3 * K / (e2e_ms * 1e-3) / 1e6, "unit": "Mcell-updates/s",
           "h2d_bytes_per_step": h2d / K + 24, "d2h_bytes_per_step": d2h / K + 16,
           "what": "pinned-host upload of u,p + K x sim_step(sim) (Python API -> C ABI, dt/n read back every step) + download of u,p; max(CUDA events, wall clock)",
           "vcycles": list(sim.pois.n[-2 * K:])}

    # ---- CPU baseline on a bounded sample --------------------------------------------------------
    cpu = None
    if world > 1:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    if not args.no_cpu and world == 1:
        cv, cms, cores, cns = cpu_oracle_run(args.cpu_size, args.cpu_steps, 1)
        cpu = {"value": cv, "unit": "Mcell-updates/s", "cores": cores, "kind": "port",
               "sample": f"TGV {args.cpu_size}^3 periodic Re=1600, {args.cpu_steps} mom_step! after 1 warm-up ({cms:.0f} ms/step, V-cycles {cns}); C++/OpenMP oracle — the Julia reference cannot run in this image"}

    line = {
        "metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(N, world),
        "details": {"multigrid_levels": eng.nlevels(), "vcycles_per_solve": ns, "planes_per_rank": info["nz_local"],
                    "nccl_calls_per_step": collectives / K if world > 1 else 0},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512, help="TGV edge length N (BASELINE.json: 512)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-size", type=int, default=256, help="edge length of the CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
