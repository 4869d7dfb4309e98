"""c4 shape (funnel d = 10, 65536 chains, auto group) for ncu captures of the persistent fused kernel."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab
from aehmc_b200 import _engine
Cn = 65536
model = ab.models.NealFunnel(10)
state = ab.nuts.new_state(np.random.default_rng(0).standard_normal((Cn, 10)), model)
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    info, ex = _engine.run("nuts", model, np.ones(10), ab.RandomStream(seed=11), state, 0.1, n_transitions=20, return_counters=True)
    e1.record(); torch.cuda.synchronize()
leap = int(ex["counters"][0])
print("ok leapfrogs", leap, "ms", e0.elapsed_time(e1), "evals/s %.3e" % (leap / e0.elapsed_time(e1) * 1e3))
