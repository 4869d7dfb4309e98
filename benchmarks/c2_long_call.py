"""c2 steady state: per-tick time when many ticks run inside ONE library call vs 24-tick calls (development tool)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab, bench
from aehmc_b200 import _engine
_, Cn, d, _, _ = bench.WORKLOADS["c2"]
cov, prec = bench.make_dense_problem(d)
EPS = bench.EPS["dense"]
dev = torch.device("cuda:0")
model = ab.models.CorrelatedGaussian(np.zeros(d), prec, device=dev)
metric = ab.metrics.GaussianMetric(cov, torch.float64, dev)
srng = ab.RandomStream(seed=2026)
state = ab.nuts.new_state(torch.from_numpy(bench.initial_positions("dense", Cn, d)).to(dev), model)
info, ex = _engine.run("nuts", model, metric, srng, state, EPS, max_ticks=96, workspace_key="k", return_counters=True)
for ticks in (24, 240, 24, 240):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_calls = 240 // ticks
    for _ in range(n_calls):
        info, ex = _engine.run("nuts", model, metric, srng, info.state, EPS, max_ticks=ticks, resume=True, workspace_key="k", return_counters=True)
    e1.record(); torch.cuda.synchronize()
    print(f"ticks/call {ticks}: {e0.elapsed_time(e1)/240:.3f} ms per tick", flush=True)
