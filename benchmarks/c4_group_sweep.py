import sys, os, json, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab
from aehmc_b200 import _engine
Cn = 65536
for name, model, eps in (("funnel", ab.models.NealFunnel(10), 0.1), ("schools", ab.models.EightSchools(), 0.39)):
    q0 = np.random.default_rng(0).standard_normal((Cn, 10))
    state = ab.nuts.new_state(q0, model)
    for G in (1, 8):
        for rep in range(2):
            srng = ab.RandomStream(seed=11)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            info, ex = _engine.run("nuts", model, np.ones(10), srng, state, eps, n_transitions=100, group=G, return_counters=True)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); leap = int(ex["counters"][0])
        print(name, "G", G, "ms", round(ms,1), "evals/s", f"{leap/ms*1e3:.3e}", "leap", leap, flush=True)
