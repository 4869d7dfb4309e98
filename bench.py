#!/usr/bin/env python
"""Throughput benchmark of the many-chain NUTS hot path (BASELINE.json metric: leapfrog gradient evals/s).

Workload (N=1): BASELINE.json configs[1] -- NUTS on a 1000-dim correlated Gaussian with a dense inverse mass
matrix, 4096 chains per GPU (synthetic inputs of SURVEY.md section 8d).  A "step" is TICKS engine ticks in
free-running mode; every tick is one velocity-Verlet step (one gradient evaluation) of every chain, with
chains finishing and restarting NUTS transitions independently.  N>1: chains are sharded (4096 per GPU,
Philox keyed by global chain id), no data-path collective ("weak" scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c2small]
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (chains per GPU, dim, ticks per step)
    "c2": (4096, 1000, 24),
    "c2small": (512, 256, 8),
}
EPS = 0.25
METRIC = "leapfrog_gradient_evals_per_sec"
UNIT = "gradient evals/s"


def make_problem(d):
    """Sigma = A A^T / d + 0.1 I (seed 3), Lambda = Sigma^-1, imm = Sigma."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.1 * np.eye(d)
    prec = np.linalg.inv(cov)
    prec = 0.5 * (prec + prec.T)
    return cov, prec


def initial_positions(C, d, chain_offset=0):
    rng = np.random.default_rng([5, chain_offset])
    return rng.standard_normal((C, d))


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy restatement of the reference, one chain per process, 1 BLAS thread each)
# ----------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    d, n_transitions, seed = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    from oracle import kernels, models, streams
    cov, prec = make_problem(d)
    model = models.CorrelatedGaussian(np.zeros(d), prec)
    srng = streams.StreamDraws(seed, "nuts")
    kernel = kernels.nuts_new_kernel(srng, model)
    state = kernels.new_state(np.random.default_rng(seed).standard_normal(d), model)
    n_leap = 0
    t0 = time.perf_counter()
    for _ in range(n_transitions):
        info, extras = kernel(state, EPS, cov)
        n_leap += extras["n_leapfrog"]
        state = info.state._replace(momentum=None)
    dt = time.perf_counter() - t0
    del limiter
    return n_leap, dt


def cpu_reference_sample(d, cores, n_transitions):
    """All host cores, one oracle chain each; returns (evals/s aggregate, leapfrogs, wall seconds)."""
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(d, n_transitions, 1000 + i) for i in range(cores)])
    wall = time.perf_counter() - t0
    n_leap = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return n_leap / busy, n_leap, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C, d, ticks = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_tr = 30 if d >= 1000 else 60
    vals = []
    warm = min(args.warmup, 1)
    for i in range(warm + args.steps):
        v, n_leap, wall = cpu_reference_sample(d, cores, n_tr)
        if i >= warm:
            vals.append((v, n_leap, wall))
    value = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[2] for v in vals]) * 1e3)
    sample = f"{cores} oracle chains (one per core, 1 BLAS thread each) x {n_tr} NUTS transitions of the {args.workload} target per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: NUTS, {d}-dim correlated Gaussian, dense inverse mass matrix "
                               f"(CPU arm: NumPy oracle restating aesara-devs/aehmc; the real reference needs Aesara, "
                               f"which is not installable here)", "chains": cores, "dim": d, "step_size": EPS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, through NVML in a background thread (a looping
    `nvidia-smi -lms` process perturbs kernel launches enough to halve the measured throughput)."""

    def __init__(self, index, period=0.1):
        self.index, self.period, self.rows, self.stop_flag, self.thread, self.ok = index, period, [], False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a list of ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((sm, mx, int(reasons)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({n for _, _, r in self.rows for n, bit in names.items() if r & bit})
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if sm else None,
                "reasons": reasons, "samples": len(sm)}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import aehmc_b200 as ab
    from aehmc_b200 import _engine, _lib, backend
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    Cn, d, ticks = WORKLOADS[args.workload]
    cov, prec = make_problem(d)
    model = ab.models.CorrelatedGaussian(np.zeros(d), prec, device=dev)
    metric = ab.metrics.GaussianMetric(cov, torch.float64, dev)
    chain_offset = rank * Cn
    q_host = torch.from_numpy(initial_positions(Cn, d, chain_offset)).pin_memory()
    srng = ab.RandomStream(seed=2026, chain_offset=chain_offset)
    key = ("bench", rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm: state lives in the engine workspace, each step continues it -----------------
    state = ab.nuts.new_state(q_host.to(dev), model)
    info, extras = _engine.run("nuts", model, metric, srng, state, EPS, max_ticks=ticks, workspace_key=key,
                               return_counters=True)
    state = info.state

    def step_resident():
        nonlocal state
        info, ex = _engine.run("nuts", model, metric, srng, state, EPS, max_ticks=ticks, resume=True,
                               workspace_key=key, return_counters=True)
        state = info.state
        return ex["counters"]

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    counters = []
    for _ in range(args.steps):
        counters.append(step_resident())
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    cnt = torch.stack(counters).sum(0).cpu().numpy()      # leapfrogs, transitions, ticks(unused), chain-ticks
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    leap = torch.tensor([float(cnt[0]), float(cnt[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(leap, op=dist.ReduceOp.SUM)
    ms_total = t.item()
    total_leapfrogs, total_transitions = leap[0].item(), leap[1].item()
    value = total_leapfrogs / (ms_total * 1e-3)

    # ---- end-to-end arm: host buffers in, host buffers out, through the public API ----------------------
    q_out = torch.empty((Cn, d), dtype=torch.float64).pin_memory()
    acc_out = torch.empty(Cn, dtype=torch.float64).pin_memory()

    def step_e2e():
        q_dev = q_host.to(dev, non_blocking=True)
        st = ab.nuts.new_state(q_dev, model)
        info, ex = _engine.run("nuts", model, metric, srng, st, EPS, max_ticks=ticks, workspace_key=key,
                               return_counters=True)
        q_out.copy_(info.state.position, non_blocking=True)
        acc_out.copy_(info.acceptance_probability, non_blocking=True)
        return ex["counters"]

    for _ in range(3):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cs = [step_e2e() for _ in range(args.steps)]
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    leap_e2e = torch.tensor([float(torch.stack(cs).sum(0)[0].item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(leap_e2e, op=dist.ReduceOp.SUM)
    e2e_value = leap_e2e.item() / (ms_e2e.item() * 1e-3)

    # ---- roofline of the dominant kernel (dense_apply_kernel<double>), timed alone on the same stream -----
    lib = _lib.load()
    a = torch.randn((Cn, d), dtype=torch.float64, device=dev)
    out = torch.empty_like(a)
    ctx = backend.context(dev)

    def gemm():
        _lib.check(lib.b2h_dense_apply(ctx, _lib.F64, backend.ptr(a), backend.ptr(metric.imm), backend.ptr(out),
                                       C.c_int64(Cn), C.c_int64(d)))
    for _ in range(3):
        gemm()
    torch.cuda.synchronize(dev)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    g0.record()
    for _ in range(reps):
        gemm()
    g1.record()
    torch.cuda.synchronize(dev)
    gemm_ms = g0.elapsed_time(g1) / reps
    flops = 2.0 * Cn * d * d
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    # FP64 peak is not in MEASURED_PEAKS.json: measure cuBLAS DGEMM here, the way the driver measured bf16
    n = 4096
    x = torch.randn((n, n), dtype=torch.float64, device=dev)
    y = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(x, y)
    best = 1e9
    for _ in range(5):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(); torch.matmul(x, y); p1.record(); torch.cuda.synchronize(dev)
        best = min(best, p0.elapsed_time(p1))
    fp64_peak = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    # the same product back to back for ~1.5 s: what the FP64 tensor path sustains under the power cap
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps_s = max(10, int(1500.0 / best))
    p0.record()
    for _ in range(reps_s):
        torch.matmul(x, y)
    p1.record(); torch.cuda.synchronize(dev)
    fp64_sustained = 2.0 * n ** 3 * reps_s / (p0.elapsed_time(p1) * 1e-3) / 1e12
    del x, y
    # cuBLAS on the kernel's own shape, for context
    mm = metric.imm
    torch.matmul(a, mm)
    best = 1e9
    for _ in range(5):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(); torch.matmul(a, mm); p1.record(); torch.cuda.synchronize(dev)
        best = min(best, p0.elapsed_time(p1))
    cublas_same_shape = flops / (best * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    gemms_per_tick = 4.0          # 3 whole-batch applies + the (small) momentum applies, upper bound on their share
    step_ms = ms_total / args.steps
    elementwise_ms = max(step_ms - ticks * 2.0 * gemm_ms, 1e-9)
    b_nuts = 11.0 * d * 8.0       # algorithmic bytes of one NUTS inner step incl. U-turn bookkeeping
    hbm_achieved = b_nuts * Cn * ticks / (elementwise_ms * 1e-3) / 1e9

    # ---- second metric of BASELINE.json: NUTS ESS/s (min over the monitored dims, all chains, all ranks) ----
    ess_per_s = None
    if not args.no_ess:
        n_tr = args.ess_transitions
        st0 = ab.nuts.new_state(q_host.to(dev), model)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        info, ex = _engine.run("nuts", model, metric, ab.RandomStream(seed=7, chain_offset=chain_offset), st0, EPS,
                               n_transitions=n_tr, store_draws=n_tr, workspace_key=("ess", rank))
        s1.record()
        barrier()
        t_ess = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ess, op=dist.ReduceOp.MAX)
        burn = n_tr // 4
        ess = ab.diagnostics.ess(ex["draws"][burn:], dims=list(range(min(8, d))))
        ess_per_s = float(np.nanmin(ess)) / (t_ess.item() * 1e-3)
        del ex

    if rank == 0:
        kernels_per_tick = 6          # pre, gradient apply, potential, imm.g apply, post + the momentum side launch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: NUTS, {d}-dim correlated Gaussian, dense inverse mass matrix, "
                                   f"{Cn} chains per GPU (BASELINE.json configs[1])", "chains_per_gpu": Cn, "dim": d,
                       "step_size": EPS, "max_num_expansions": 10, "ticks_per_step": ticks, "rng": "philox4x32-10",
                       "l2": "engine state ~1.8 GB per GPU, larger than the 126 MB L2 (no flush needed)",
                       "transitions_per_step": total_transitions / args.steps,
                       "mean_leapfrogs_per_transition": total_leapfrogs / max(total_transitions, 1.0)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(Cn * d * 8),
                    "d2h_bytes_per_step": int(Cn * d * 8 + Cn * 8),
                    "what": "pinned host positions -> device, new_state, TICKS ticks, position + acceptance back to pinned host"},
            "gpu_launches": int(args.steps * (ticks * kernels_per_tick + 3)),
            "clocks": clocks,
            "nuts_ess_per_sec": ess_per_s,
            "ess_how": None if ess_per_s is None else
            f"{args.ess_transitions} NUTS transitions per chain from the initial positions, first quarter discarded, "
            "multi-chain ESS (Stan/arviz estimator, no rank normalisation) of the first 8 coordinates, minimum, "
            "divided by the wall time of all transitions incl. the discarded ones",
            "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": None,
                         "kernel": "dense_apply_dmma_async_kernel<BN> (out[C x d] = in[C x d] . M[d x d], FP64 DMMA + cp.async): gradient and imm.g, 2 launches per tick",
                         "flops_per_launch": flops, "avg_launch_ms": gemm_ms,
                         "peak_source": f"measured in this run: torch.matmul fp64 {n}^3 (cuBLAS), best of 5 "
                                        "(MEASURED_PEAKS.json has no FP64 figure; SURVEY.md 8d names FP64 FMA as the bound)",
                         "peak_sustained": fp64_sustained, "frac_of_sustained": achieved / fp64_sustained,
                         "launches_per_tick": 2, "share_of_step": ticks * 2.0 * gemm_ms / step_ms,
                         "cublas_same_shape_tflops": cublas_same_shape},
            "roofline_elementwise": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": hbm_achieved / hbm_peak, "traffic": None,
                                     "how": "derived: 11*d*8 algorithmic bytes per chain-tick over (step time - 2 dense applies per tick); pre + post + potential kernels"},
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n_tr = 30 if d >= 1000 else 60
            v, n_leap, wall = cpu_reference_sample(d, cores, n_tr)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{cores} oracle chains (one per core) x {n_tr} NUTS transitions, "
                                              f"{n_leap} leapfrogs in {wall:.1f} s wall"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ess", action="store_true", help="skip the ESS/s leg")
    ap.add_argument("--ess-transitions", type=int, default=40)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
