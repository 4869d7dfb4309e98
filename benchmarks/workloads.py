#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 0, 2, 3 and the stand-alone integrator), one JSON line each.
bench.py stays the driver-facing benchmark (config 1 shape = c2); this script fills the rest of the
BASELINE.md table.   python benchmarks/workloads.py [c1] [c3] [c4] [leapfrog] [--small]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aehmc_b200 as ab  # noqa: E402
from aehmc_b200 import _engine, _lib, backend  # noqa: E402

PEAK_HBM = 6545.0
try:
    PEAK_HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def c1(small):
    """config 0: HMC, velocity_verlet, L = 10, diagonal imm, 100-dim iid Gaussian (eps = 0.25)."""
    d, L = 100, 10
    sigma = np.exp(0.5 * np.random.default_rng(1).standard_normal(d))
    model = ab.models.IIDGaussian(np.zeros(d), sigma)
    out = []
    for Cn in ((1, 4096) if small else (1, 65536, 1 << 20)):
        q0 = np.random.default_rng(2).standard_normal((Cn, d))
        kernel = ab.hmc.new_kernel(ab.RandomStream(seed=1), model)
        state = ab.hmc.new_state(q0, model)
        n_tr = 1000 if Cn == 1 else (100 if Cn <= 65536 else 20)
        run = lambda: _engine.run("hmc", model, sigma ** 2, kernel.spec["srng"], state, 0.25, n_transitions=n_tr,
                                  num_integration_steps=L, store_draws=0)
        run()
        (info, ex), ms = timed(run)
        evals = Cn * n_tr * L
        out.append({"workload": "c1 HMC L=10 d=100 iid Gaussian diag imm", "chains": Cn, "transitions": n_tr,
                    "grad_evals_per_sec": evals / (ms * 1e-3), "ms": ms,
                    "mean_accept": float(info.acceptance_probability.mean()),
                    "hbm_frac_6ds": evals * 6 * d * 8 / (ms * 1e-3) / 1e9 / PEAK_HBM})
    return out


def leapfrog(small):
    """Stand-alone fused integrator (b2h_leapfrog): 6*d*s algorithmic bytes per leapfrog."""
    out = []
    for dt, s in ((torch.float64, 8), (torch.float32, 4)):
        Cn, d = (65536, 128) if small else (1 << 20, 128)
        rng = np.random.default_rng(0)
        sigma = np.exp(0.3 * rng.standard_normal(d))
        model = ab.models.IIDGaussian(np.zeros(d), sigma, dtype=dt)
        _, kin, _ = ab.metrics.gaussian_metric(sigma ** 2, dtype=dt)
        lib = _lib.load()
        q = torch.randn((Cn, d), dtype=dt, device="cuda"); p = torch.randn_like(q)
        U, g = model.potential_and_grad(q)
        eps = torch.full((Cn,), 0.1, dtype=torch.float64, device="cuda")
        m, mt = model.struct(), kin.metric.struct()
        ctx = backend.context(q.device)
        for n_steps in (1, 8):
            fn = lambda: _lib.check(lib.b2h_leapfrog(ctx, C.byref(m), C.byref(mt), backend.code(dt), backend.ptr(q),
                                                     backend.ptr(p), backend.ptr(U), backend.ptr(g), backend.ptr(eps), None,
                                                     C.c_int32(n_steps), C.c_int64(Cn), None, C.c_int64(0)))
            for _ in range(3):
                fn()
            _, ms = timed(fn, 20)
            bytes_launch = 6 * d * s * Cn                      # one read + one write of q, p, g per launch
            out.append({"workload": f"b2h_leapfrog iid Gaussian d={d} {str(dt)[6:]}", "chains": Cn, "n_steps": n_steps,
                        "ms": ms, "leapfrogs_per_sec": Cn * n_steps / (ms * 1e-3),
                        "achieved_GBs_launch_traffic": bytes_launch / (ms * 1e-3) / 1e9,
                        "hbm_frac": bytes_launch / (ms * 1e-3) / 1e9 / PEAK_HBM,
                        "algorithmic_frac_6ds_per_leapfrog": 6 * d * s * Cn * n_steps / (ms * 1e-3) / 1e9 / PEAK_HBM})
    return out


def uturn(small):
    """Stand-alone metric / U-turn primitives (b2h_kinetic_energy, b2h_is_turning, b2h_termination_update,
    b2h_is_iterative_turning): achieved GB/s by their algorithmic bytes (streaming formulation)."""
    out = []
    lib = _lib.load()
    for dt, s in ((torch.float64, 8), (torch.float32, 4)):
        Cn, d, maxd = (65536, 128, 10) if small else (262144, 128, 10)
        code = backend.code(dt)
        imm = torch.rand(d, dtype=torch.float64).add(0.5).numpy()
        metric = ab.metrics.GaussianMetric(imm, dt, torch.device("cuda"))
        mt = metric.struct()
        ctx = backend.context(torch.device("cuda"))
        g = torch.Generator(device="cuda").manual_seed(1)
        rnd = lambda *shape: torch.randn(shape, dtype=dt, device="cuda", generator=g)
        p, pl, pr, ps = rnd(Cn, d), rnd(Cn, d), rnd(Cn, d), rnd(Cn, d)
        mck, sck = rnd(Cn, maxd, d), rnd(Cn, maxd, d)
        K = torch.empty(Cn, dtype=dt, device="cuda")
        flag = torch.empty(Cn, dtype=torch.uint8, device="cuda")
        imin = torch.zeros(Cn, dtype=torch.int64, device="cuda")
        imax = torch.zeros(Cn, dtype=torch.int64, device="cuda")
        step = torch.full((Cn,), 6, dtype=torch.int64, device="cuda")            # even step: one checkpoint row written
        n64 = C.c_int64
        calls = {
            "kinetic_energy": (lambda: lib.b2h_kinetic_energy(ctx, C.byref(mt), code, backend.ptr(p), backend.ptr(K), n64(Cn),
                                                              n64(d), None, n64(0)), 1 * d * s),
            "is_turning": (lambda: lib.b2h_is_turning(ctx, C.byref(mt), code, backend.ptr(pl), backend.ptr(pr), backend.ptr(ps),
                                                      backend.ptr(flag), n64(Cn), n64(d), None, n64(0)), 3 * d * s),
            "termination_update (even step)": (lambda: lib.b2h_termination_update(
                ctx, code, backend.ptr(mck), backend.ptr(sck), backend.ptr(imin), backend.ptr(imax), backend.ptr(ps),
                backend.ptr(p), backend.ptr(step), n64(Cn), n64(d), C.c_int32(maxd)), 4 * d * s),
        }
        for levels in (1, 3):
            imin.zero_(); imax.fill_(levels - 1)
            lo, hi = imin.clone(), imax.clone()
            # make the criterion hold at every level so that all `levels` rows are read: msum far along +p
            big = (p * 1000.0).contiguous()
            mck.copy_(p.unsqueeze(1).expand(Cn, maxd, d)); sck.zero_()
            calls[f"is_iterative_turning ({levels} level{'s' if levels > 1 else ''})"] = (
                (lambda lo=lo, hi=hi, big=big: lib.b2h_is_iterative_turning(
                    ctx, C.byref(mt), code, backend.ptr(mck), backend.ptr(sck), backend.ptr(lo), backend.ptr(hi),
                    backend.ptr(big), backend.ptr(p), backend.ptr(flag), n64(Cn), n64(d), C.c_int32(maxd))),
                (2 + 2 * levels) * d * s)
        for name, (fn, bytes_chain) in calls.items():
            for _ in range(3):
                _lib.check(fn())
            _, ms = timed(lambda: _lib.check(fn()), 20)
            gbs = bytes_chain * Cn / (ms * 1e-3) / 1e9
            out.append({"workload": f"{name} d={d} {str(dt)[6:]}", "chains": Cn, "ms": ms, "achieved_GBs": gbs,
                        "hbm_frac": gbs / PEAK_HBM, "algorithmic_bytes_per_chain": bytes_chain})
    return out


def c4(small):
    """config 3: window_adaptation (1000 steps) + 1000 NUTS draws, 10-dim funnel and eight schools, 65536 chains."""
    out = []
    Cn = 4096 if small else 65536
    W = D = 200 if small else 1000
    for name, model in (("funnel", ab.models.NealFunnel(10)), ("eight_schools", ab.models.EightSchools())):
        q0 = np.random.default_rng(0).standard_normal((Cn, 10))
        kernel = ab.nuts.new_kernel(ab.RandomStream(seed=11), model)
        state = ab.nuts.new_state(q0, model)
        t0 = time.perf_counter()
        (wstate, (eps, imm), upd), ms_w = timed(lambda: ab.window_adaptation.run(kernel, state, W))
        from aehmc_b200 import metrics
        run = lambda: _engine.run("nuts", model, metrics.per_chain(imm), kernel.spec["srng"], wstate, eps,
                                  n_transitions=D, store_draws=0, return_counters=True)
        (info, ex), ms_d = timed(run)
        leap = int(ex["counters"][0].item())
        depth = torch.bincount(info.num_doublings.long(), minlength=11).cpu().numpy().tolist()
        out.append({"workload": f"c4 window_adaptation({W}) + {D} NUTS draws, {name} d=10", "chains": Cn,
                    "warmup_ms": ms_w, "sampling_ms": ms_d, "sampling_leapfrogs": leap,
                    "grad_evals_per_sec": leap / (ms_d * 1e-3),
                    "nuts_bytes_frac_11ds": leap * 11 * 10 * 8 / (ms_d * 1e-3) / 1e9 / PEAK_HBM,
                    "transitions_per_sec": Cn * D / (ms_d * 1e-3),
                    "step_size_median": float(eps.median()), "last_transition_depth_hist": depth,
                    "last_transition_divergent_frac": float(info.is_diverging.double().mean())})
    return out


def wide(small):
    """Fused persistent NUTS kernel on an elementwise target (iid Gaussian, diagonal imm): the streaming
    formulation's 11*d*s algorithmic bytes per leapfrog against the HBM peak."""
    out = []
    for d, Cn, dt, G in ((128, 131072, torch.float64, 0), (128, 131072, torch.float32, 0), (1000, 16384, torch.float64, 0),
                         (32, 262144, torch.float64, 0)):
        if small:
            Cn //= 8
        rng = np.random.default_rng(0)
        sigma = np.exp(0.3 * rng.standard_normal(d))
        model = ab.models.IIDGaussian(np.zeros(d), sigma, dtype=dt)
        q0 = rng.standard_normal((Cn, d)) * sigma
        state = ab.nuts.new_state(q0, model)
        eps = 1.2 / d ** 0.25
        run = lambda: _engine.run("nuts", model, sigma ** 2, ab.RandomStream(seed=5), state, eps, n_transitions=20,
                                  return_counters=True, group=G)
        run()
        (info, ex), ms = timed(run)
        leap = int(ex["counters"][0].item())
        s_ = 8 if dt == torch.float64 else 4
        out.append({"workload": f"fused NUTS iid Gaussian d={d} {str(dt)[6:]}", "chains": Cn, "transitions": 20,
                    "ms": ms, "grad_evals_per_sec": leap / (ms * 1e-3), "mean_leapfrogs": leap / (Cn * 20),
                    "hbm_frac_11ds": leap * 11 * d * s_ / (ms * 1e-3) / 1e9 / PEAK_HBM,
                    "mean_accept": float(info.acceptance_probability.mean())})
    return out


def c3(small):
    """config 2: NUTS Bayesian logistic regression, N = 100k, D = 128, 4096 chains (FP64/FP32 FMA-path gradient)."""
    out = []
    N, D, Cn = (20000, 128, 1024) if small else (100000, 128, 4096)
    rng = np.random.default_rng(4)
    X = rng.standard_normal((N, D)).astype(np.float32)
    X = torch.tensor(X).bfloat16().double().numpy()              # bf16-representable
    beta = rng.standard_normal(D) / np.sqrt(D)
    y = (rng.random(N) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    q0 = 0.1 * np.random.default_rng(6).standard_normal((Cn, D))
    cases = ((torch.float64, False), (torch.float32, False), (torch.float32, "two_kernel"), (torch.float32, "bf16x3"),
             (torch.float32, True), (torch.float64, True))
    if "--tc-only" in sys.argv:
        cases = cases[2:]
    for dt, tcore in cases:
        model = ab.models.LogisticRegression(X, y, 1.0, dtype=dt, tensor_core=tcore)
        imm = np.full(D, 4.0 / N)
        srng = ab.RandomStream(seed=3)
        state = ab.nuts.new_state(q0, model)
        ticks = 6
        run = lambda: _engine.run("nuts", model, imm, srng, state, 0.4, max_ticks=ticks, return_counters=True)
        run()
        (info, ex), ms = timed(run)
        leap = int(ex["counters"][0].item())
        flops = 4.0 * N * D * leap
        pieces = 2 if getattr(model, "tc_flag", 0.0) == 4.0 else 3
        path = {0.0: "FMA/DMMA-path gradient", 2.0: "tcgen05 bf16x3 gradient, fully fused", 3.0: "tcgen05 bf16x3 gradient, two kernels",
                4.0: "tcgen05 fp16x2 gradient, fully fused, A operands in TMEM"}[getattr(model, "tc_flag", 0.0)]
        out.append({"workload": f"c3 NUTS logistic N={N} D={D} {str(dt)[6:]} ({path})", "chains": Cn,
                    "ticks": ticks, "ms_per_tick": ms / ticks, "grad_evals_per_sec": leap / (ms * 1e-3),
                    "gradient_TFLOPs_algorithmic_4ND": flops / (ms * 1e-3) / 1e12,
                    "tensor_TFLOPs_issued": (pieces * flops / (ms * 1e-3) / 1e12) if tcore else None,
                    "mean_accept": float(info.acceptance_probability.mean())})
    return out


if __name__ == "__main__":
    small = "--small" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["leapfrog", "c1", "wide", "c4", "c3"]
    for n in names:
        for line in {"c1": c1, "c3": c3, "c4": c4, "leapfrog": leapfrog, "wide": wide, "uturn": uturn}[n](small):
            print(json.dumps(line), flush=True)
