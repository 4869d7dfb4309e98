"""Batched logistic-regression gradient alone (c3 shape by default): ms per evaluation of all chains, issued
tensor TFLOP/s, SM clock under load.   python benchmarks/logistic_grad.py [--chains C] [--reps R] [--two-kernel]"""
import argparse, json, os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=4096)
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--two-kernel", action="store_true")
ap.add_argument("--bf16x3", action="store_true")
ap.add_argument("--f64", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(4)
X = torch.tensor(rng.standard_normal((a.n, a.d)), dtype=torch.float32).bfloat16().double().numpy()
beta = rng.standard_normal(a.d) / np.sqrt(a.d)
y = (rng.random(a.n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
dt = torch.float64 if a.f64 else torch.float32
model = ab.models.LogisticRegression(X, y, 1.0, dtype=dt, tensor_core="two_kernel" if a.two_kernel else ("bf16x3" if a.bf16x3 else True))
q = torch.tensor(0.1 * np.random.default_rng(6).standard_normal((a.chains, a.d)), dtype=dt, device="cuda")
for _ in range(5):
    model.potential_and_grad(q)
torch.cuda.synchronize()
clocks, stop = [], False
def sample():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.01)
    except Exception:
        pass
th = threading.Thread(target=sample, daemon=True); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    model.potential_and_grad(q)
e1.record(); torch.cuda.synchronize()
stop = True; th.join(timeout=1)
ms = e0.elapsed_time(e1) / a.reps
flops = 4.0 * a.n * a.d * a.chains
print(json.dumps({"workload": f"logistic gradient N={a.n} D={a.d} chains={a.chains} {'two-kernel' if a.two_kernel else 'fused'} path={model.tc_flag} "
                  , "ms_per_gradient": ms,
                  "evals_per_sec": a.chains / (ms * 1e-3), "TFLOPs_algorithmic": flops / (ms * 1e-3) / 1e12,
                  "TFLOPs_issued": (2 if model.tc_flag == 4.0 else 3) * flops / (ms * 1e-3) / 1e12,
                  "sm_mhz_median": float(np.median(clocks)) if clocks else None, "sm_mhz_min": min(clocks) if clocks else None}))
