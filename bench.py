#!/usr/bin/env python
"""Throughput benchmark of the many-chain NUTS hot path (BASELINE.json metric: leapfrog gradient evals/s).

Default workload (N=1): BASELINE.json configs[1] ("c2") -- NUTS on a 1000-dim correlated Gaussian with a dense
inverse mass matrix, 4096 chains per GPU (synthetic inputs of SURVEY.md section 8d).  Other workloads:
"c3" = configs[2] (NUTS Bayesian logistic regression, N = 100k, D = 128, 4096 chains, tcgen05 gradient) and
"c5" = configs[4] (the same model with 131072 chains per GPU = 1M chains on 8 GPUs).  A "step" is TICKS engine
ticks in free-running mode; every tick is one velocity-Verlet step (one gradient evaluation) of every chain,
with chains finishing and restarting NUTS transitions independently.  N>1: chains are sharded (Philox keyed by
global chain id), no data-path collective ("weak" scaling); the only collective is the all-reduce of the
R-hat / ESS sufficient statistics.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c5|c2small|c3small]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, chains per GPU, dim, ticks per step, data rows)
    "c2": ("dense", 4096, 1000, 24, 0),
    "c2small": ("dense", 512, 256, 8, 0),
    "c3": ("logistic", 4096, 128, 24, 100000),
    "c5": ("logistic", 131072, 128, 8, 100000),
    "c3small": ("logistic", 512, 64, 8, 4096),
}
EPS = {"dense": 0.25, "logistic": 0.4}
METRIC = "leapfrog_gradient_evals_per_sec"
UNIT = "gradient evals/s"


# ----------------------------------------------------------------------------------------------------------
# synthetic problems; NumPy only, shared by the GPU arm and the CPU arm
# ----------------------------------------------------------------------------------------------------------
def make_dense_problem(d):
    """config 2: Sigma = A A^T / d + 0.1 I (seed 3), Lambda = Sigma^-1, imm = Sigma."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.1 * np.eye(d)
    prec = np.linalg.inv(cov)
    prec = 0.5 * (prec + prec.T)
    return cov, prec


def bf16_round(x):
    """Round float32 values to the nearest bf16-representable value (ties to even), in NumPy."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.view(np.float32)


def make_logistic_problem(n, d):
    """configs 3/5: X ~ N(0,1) rounded to bf16 (seed 4), beta* ~ N(0,1)/sqrt(D), y ~ Bernoulli(sigmoid(X beta*)),
    prior scale 1, diagonal imm = 4/N."""
    rng = np.random.default_rng(4)
    X = bf16_round(rng.standard_normal((n, d)).astype(np.float32)).astype(np.float64)
    beta = rng.standard_normal(d) / np.sqrt(d)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    return X, y, np.full(d, 4.0 / n)


def initial_positions(kind, C, d, chain_offset=0):
    if kind == "dense":
        return np.random.default_rng([5, chain_offset]).standard_normal((C, d))
    return 0.1 * np.random.default_rng([6, chain_offset]).standard_normal((C, d))


def describe(name):
    kind, Cn, d, ticks, n = WORKLOADS[name]
    if kind == "dense":
        which = " (BASELINE.json configs[1])" if name == "c2" else ""
        return f"{name}: NUTS, {d}-dim correlated Gaussian, dense inverse mass matrix, {Cn} chains per GPU{which}"
    which = {"c3": " (BASELINE.json configs[2])", "c5": " (BASELINE.json configs[4]: 1M chains on 8 GPUs)"}.get(name, "")
    return (f"{name}: NUTS, Bayesian logistic regression N={n} D={d}, diagonal inverse mass matrix, "
            f"{Cn} chains per GPU{which}")


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy restatement of the reference, one chain per process, 1 BLAS thread each)
# ----------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    name, n_transitions, seed = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    from oracle import kernels, models, streams
    kind, _, d, _, n = WORKLOADS[name]
    if kind == "dense":
        cov, prec = make_dense_problem(d)
        model, imm = models.CorrelatedGaussian(np.zeros(d), prec), cov
    else:
        X, y, imm = make_logistic_problem(n, d)
        model = models.LogisticRegression(X, y, 1.0)
    srng = streams.StreamDraws(seed, "nuts")
    kernel = kernels.nuts_new_kernel(srng, model)
    q0 = initial_positions(kind, 1, d, seed)[0]
    state = kernels.new_state(q0, model)
    n_leap = 0
    t0 = time.perf_counter()
    for _ in range(n_transitions):
        info, extras = kernel(state, EPS[kind], imm)
        n_leap += extras["n_leapfrog"]
        state = info.state._replace(momentum=None)
    dt = time.perf_counter() - t0
    del limiter
    return n_leap, dt


def cpu_transitions(name, n_samples=1):
    """NUTS transitions per oracle chain in one CPU sample; shrunk when many samples are requested so that the
    whole reference arm stays within a few minutes (about 0.3 s per c2 transition, 0.2 s per c3 transition)."""
    kind, _, d, _, n = WORKLOADS[name]
    if kind == "dense":
        full = 30 if d >= 1000 else 60
        return max(4, min(full, (8 * full) // max(n_samples, 1)))
    full = 20 if n >= 100000 else 60
    return max(3, min(full, (8 * full) // max(n_samples, 1)))


def cpu_reference_sample(name, cores, n_transitions):
    """All host cores, one oracle chain each; returns (evals/s aggregate, leapfrogs, wall seconds)."""
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(name, n_transitions, 1000 + i) for i in range(cores)])
    wall = time.perf_counter() - t0
    n_leap = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return n_leap / busy, n_leap, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, _, d, _, n = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    warm = min(args.warmup, 1)
    n_tr = cpu_transitions(args.workload, warm + args.steps)
    vals = []
    for i in range(warm + args.steps):
        v, n_leap, wall = cpu_reference_sample(args.workload, cores, n_tr)
        if i >= warm:
            vals.append((v, n_leap, wall))
    value = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[2] for v in vals]) * 1e3)
    sample = (f"{cores} oracle chains (one per core, 1 BLAS thread each) x {n_tr} NUTS transitions of the "
              f"{args.workload} target per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(args.workload) + " -- CPU arm: NumPy oracle restating aesara-devs/aehmc "
                               "(the real reference needs Aesara, which is not installable here)",
                   "chains": cores, "dim": d, "step_size": EPS[kind]},
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

    def __init__(self, index, period=0.02):
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


def _event_ms(fn, reps, dev):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def dense_roofline(metric, Cn, d, ticks, step_ms, hbm_peak, dev):
    """Dominant kernel of c2: the FP64 dense apply (DMMA tensor path), timed alone on the engine's stream."""
    import ctypes as C
    import torch
    from aehmc_b200 import _lib, backend
    lib = _lib.load()
    a = torch.randn((Cn, d), dtype=torch.float64, device=dev)
    out = torch.empty_like(a)
    ctx = backend.context(dev)

    def gemm():
        _lib.check(lib.b2h_dense_apply(ctx, _lib.F64, backend.ptr(a), backend.ptr(metric.imm), backend.ptr(out),
                                       C.c_int64(Cn), C.c_int64(d)))
    for _ in range(3):
        gemm()
    gemm_ms = _event_ms(gemm, 20, dev)
    flops = 2.0 * Cn * d * d
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    # FP64 peak is not in MEASURED_PEAKS.json: measure cuBLAS DGEMM here, the way the driver measured bf16
    n = 4096
    x = torch.randn((n, n), dtype=torch.float64, device=dev)
    y = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(x, y)
    best = min(_event_ms(lambda: torch.matmul(x, y), 1, dev) for _ in range(5))
    fp64_peak = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    # the same product back to back for ~1.5 s: what the FP64 tensor path sustains under the power cap
    reps_s = max(10, int(1500.0 / best))
    fp64_sustained = 2.0 * n ** 3 / (_event_ms(lambda: torch.matmul(x, y), reps_s, dev) * 1e-3) / 1e12
    del x, y
    mm = metric.imm
    torch.matmul(a, mm)
    best = min(_event_ms(lambda: torch.matmul(a, mm), 1, dev) for _ in range(5))
    cublas_same_shape = flops / (best * 1e-3) / 1e12
    elementwise_ms = max(step_ms - ticks * 2.0 * gemm_ms, 1e-9)
    b_nuts = 11.0 * d * 8.0       # algorithmic bytes of one NUTS inner step incl. U-turn bookkeeping
    hbm_achieved = b_nuts * Cn * ticks / (elementwise_ms * 1e-3) / 1e9
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r01_ncu_gemm_summary.md (ncu --set full)
        "traffic": 40.822016e6 + 6.891264e6 if (Cn, d) == (4096, 1000) else None,
        "pipe": "FP64 tensor (DMMA.8x8x4): FP64 has no tcgen05 kind",
        "kernel": "dense_apply_dmma_async_kernel<BN> (out[C x d] = in[C x d] . M[d x d], FP64 DMMA + cp.async): "
                  "gradient and imm.g, 2 launches per tick",
        "flops_per_launch": flops, "algorithmic_bytes_per_launch": 8.0 * (2 * Cn * d + d * d), "avg_launch_ms": gemm_ms,
        "peak_source": f"measured in this run: torch.matmul fp64 {n}^3 (cuBLAS), best of 5 (MEASURED_PEAKS.json "
                       "has no FP64 figure; SURVEY.md 8d names FP64 compute as the bound of config 2)",
        "peak_sustained": fp64_sustained, "frac_of_sustained": achieved / fp64_sustained,
        "launches_per_tick": 2, "share_of_step": ticks * 2.0 * gemm_ms / step_ms,
        "cublas_same_shape_tflops": cublas_same_shape}
    elementwise = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                   "frac": hbm_achieved / hbm_peak, "traffic": None,
                   "how": "derived: 11*d*8 algorithmic bytes per chain-tick over (step time - 2 dense applies per "
                          "tick); post+pre + potential kernels"}
    return roofline, elementwise


def logistic_roofline(model, Cn, d, n, ticks, step_ms, peaks, dev, dtype):
    """Dominant kernel of c3 / c5: the fused tcgen05 gradient (S product, residual, X^T R product in one kernel),
    timed through b2h_potential_and_grad on the engine's stream (includes three small side kernels)."""
    import torch
    q = torch.tensor(initial_positions("logistic", Cn, d, 12345), dtype=dtype, device=dev)
    for _ in range(3):
        model.potential_and_grad(q)
    ms = _event_ms(lambda: model.potential_and_grad(q), 20 if Cn <= 8192 else 5, dev)
    flops = 4.0 * n * d * Cn
    achieved = flops / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1389.0))
    pieces = 2 if getattr(model, "tc_flag", 0.0) == 4.0 else 3
    return {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 4096 chains (ncu --set full,
        # profiles/r01_ncu_tc_fused_summary.md): X once + the responses; everything else stays on chip
        "traffic": 28.12e6 if (Cn, d, n) == (4096, 128, 100000) else None,
        "kernel": ("tc_logistic_fused16_kernel" if pieces == 2 else "tc_logistic_fused_kernel") +
                  " (tcgen05.mma kind::f16 with both A operands in TMEM, TMA, one launch per tick): S = B X^T, "
                  "residual epilogue back into TMEM, G += R X",
        "flops_per_launch": flops, "avg_launch_ms": ms,
        "what": f"ALGORITHMIC flops (4 N D per chain-gradient).  beta and the residual are carried as {pieces} "
                f"{'fp16' if pieces == 2 else 'bf16'} pieces for fp32-class accuracy, so the tensor pipe issues "
                f"{pieces}x these flops; timed through b2h_potential_and_grad (includes 3 small side kernels)",
        "issued_tflops": pieces * achieved, "issued_frac": pieces * achieved / peak,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
        "launches_per_tick": 1, "share_of_step": ticks * ms / step_ms}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import aehmc_b200 as ab
    from aehmc_b200 import _engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kind, Cn, d, ticks, n_data = WORKLOADS[args.workload]
    eps = EPS[kind]
    if kind == "dense":
        dtype, dtype_name = torch.float64, "f64"
        cov, prec = make_dense_problem(d)
        model = ab.models.CorrelatedGaussian(np.zeros(d), prec, device=dev)
        metric = ab.metrics.GaussianMetric(cov, dtype, dev)
    else:
        dtype, dtype_name = torch.float32, "f32"
        X, y, imm = make_logistic_problem(n_data, d)
        model = ab.models.LogisticRegression(X, y, 1.0, dtype=dtype, device=dev, tensor_core=True)
        metric = ab.metrics.GaussianMetric(imm, dtype, dev)
    chain_offset = rank * Cn
    q_host = torch.from_numpy(initial_positions(kind, Cn, d, chain_offset)).to(dtype).pin_memory()
    srng = ab.RandomStream(seed=2026, chain_offset=chain_offset)
    key = ("bench", rank)
    esize = q_host.element_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm: state lives in the engine workspace, each step continues it -----------------
    state = ab.nuts.new_state(q_host.to(dev), model)
    info, extras = _engine.run("nuts", model, metric, srng, state, eps, max_ticks=ticks, workspace_key=key,
                               return_counters=True)
    state = info.state

    def step_resident():
        nonlocal state
        info, ex = _engine.run("nuts", model, metric, srng, state, eps, max_ticks=ticks, resume=True,
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
    q_out = torch.empty((Cn, d), dtype=dtype).pin_memory()
    acc_out = torch.empty(Cn, dtype=torch.float64).pin_memory()

    def step_e2e():
        q_dev = q_host.to(dev, non_blocking=True)
        st = ab.nuts.new_state(q_dev, model)
        info, ex = _engine.run("nuts", model, metric, srng, st, eps, max_ticks=ticks, workspace_key=key,
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

    # ---- roofline of the dominant kernel, timed alone on the same stream ---------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    step_ms = ms_total / args.steps
    roofline_elementwise = None
    if kind == "dense":
        roofline, roofline_elementwise = dense_roofline(metric, Cn, d, ticks, step_ms, hbm_peak, dev)
        kernels_per_tick = 6          # post+pre, gradient apply, potential, imm.g apply, momentum rider GEMM + its reduce
    else:
        roofline = logistic_roofline(model, Cn, d, n_data, ticks, step_ms, peaks, dev, dtype)
        kernels_per_tick = 5          # post+pre, beta split, response convert, fused gradient, finish

    # ---- second metric of BASELINE.json: NUTS ESS/s (min over the monitored dims, all chains, all ranks) ----
    ess_per_s = rhat_max = gathered = None
    if not args.no_ess:
        n_tr = args.ess_transitions
        st0 = ab.nuts.new_state(q_host.to(dev), model)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        info, ex = _engine.run("nuts", model, metric, ab.RandomStream(seed=7, chain_offset=chain_offset), st0, eps,
                               n_transitions=n_tr, store_draws=n_tr, workspace_key=("ess", rank))
        s1.record()
        barrier()
        t_ess = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ess, op=dist.ReduceOp.MAX)
        burn = n_tr // 4
        dims = list(range(min(8, d)))
        # sufficient statistics per rank, summed over ranks with one all-reduce (NCCL): the diagnostics gather
        ess = ab.diagnostics.ess(ex["draws"][burn:], dims=dims)
        rhat = ab.diagnostics.rhat(ex["draws"][burn:], dims=dims)
        ess_per_s = float(np.nanmin(ess)) / (t_ess.item() * 1e-3)
        rhat_max = float(np.nanmax(rhat))
        gathered = None
        if world > 1:
            # configs[4] words it as a "gather of draws": all-gather the monitored coordinates (NCCL) and recompute
            # R-hat from the gathered tensor; it must agree with the all-reduced sufficient statistics
            g = ab.diagnostics.gather_draws(ex["draws"][burn:], dims=dims)
            r2 = ab.diagnostics.rhat(g, distributed=False)
            gathered = {"shape": list(g.shape), "bytes": int(g.numel() * g.element_size()),
                        "rhat_max": float(np.nanmax(r2))}
            del g
        del ex

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": describe(args.workload), "chains_per_gpu": Cn, "chains_total": Cn * world, "dim": d,
                       "step_size": eps, "max_num_expansions": 10, "ticks_per_step": ticks, "rng": "philox4x32-10",
                       "l2": "engine state larger than the 126 MB L2 (no flush needed)" if kind == "dense" else
                             "the design matrix X (25.6 MB bf16) is meant to stay L2-resident; the chain state of "
                             "c5 (131072 x 128 x ~20 arrays) exceeds L2",
                       "transitions_per_step": total_transitions / args.steps,
                       "mean_leapfrogs_per_transition": total_leapfrogs / max(total_transitions, 1.0)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(Cn * d * esize),
                    "d2h_bytes_per_step": int(Cn * d * esize + Cn * 8),
                    "what": "pinned host positions -> device, new_state, TICKS ticks, position + acceptance back to pinned host"},
            "gpu_launches": int(args.steps * (ticks * kernels_per_tick + 3)),
            "clocks": clocks,
            "nuts_ess_per_sec": ess_per_s, "rhat_max": rhat_max, "gathered_draws": gathered if not args.no_ess else None,
            "ess_how": None if ess_per_s is None else
            f"{args.ess_transitions} NUTS transitions per chain from the initial positions, first quarter discarded, "
            "multi-chain ESS (Stan/arviz estimator, no rank normalisation) of the first 8 coordinates, minimum, "
            "divided by the wall time of all transitions incl. the discarded ones; statistics all-reduced over ranks",
            "roofline": roofline,
        }
        if roofline_elementwise is not None:
            line["roofline_elementwise"] = roofline_elementwise
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n_tr = cpu_transitions(args.workload)
            v, n_leap, wall = cpu_reference_sample(args.workload, cores, n_tr)
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
