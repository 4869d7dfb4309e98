"""c4 (funnel / eight schools, 65536 adapted chains): is the n_transitions run bound by throughput or by its slowest
chains?  Per-chain leapfrog totals (mean / quantiles / max), the time of the n_transitions run, and the time of a
free-running run of the same mean length (every lane busy all the time)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab  # noqa: E402
from aehmc_b200 import _engine, metrics  # noqa: E402

Cn, W, D = 65536, 1000, 200


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


for name, model in (("funnel", ab.models.NealFunnel(10)), ("eight_schools", ab.models.EightSchools())):
    q0 = np.random.default_rng(0).standard_normal((Cn, 10))
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=11), model)
    (wstate, (eps, imm), _), ms_w = timed(lambda: ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W))
    for G in (8, 1):
        (info, ex), ms = timed(lambda: _engine.run("nuts", model, metrics.per_chain(imm), ab.RandomStream(seed=5), wstate, eps,
                                                   n_transitions=D, store_draws=D, return_counters=True, group=G))
        per_chain = ex["draw_stats"][:, :, 2].sum(0)
        leap = float(per_chain.sum())
        q = np.quantile(per_chain.cpu().numpy(), [0.5, 0.9, 0.99, 0.999])
        print(f"{name} G={G}: {D} transitions {ms:.1f} ms, {leap / ms * 1e3:.3e} evals/s; per-chain leapfrogs mean "
              f"{float(per_chain.mean()):.0f} median {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} max {float(per_chain.max()):.0f}",
              flush=True)
        ticks = int(per_chain.mean())
        (info2, ex2), ms2 = timed(lambda: _engine.run("nuts", model, metrics.per_chain(imm), ab.RandomStream(seed=5), wstate, eps,
                                                      max_ticks=ticks, return_counters=True, group=G))
        leap2 = float(ex2["counters"][0])
        print(f"{name} G={G}: free-running {ticks} ticks {ms2:.1f} ms, {leap2 / ms2 * 1e3:.3e} evals/s", flush=True)
        del ex, ex2
