"""c4 shape, thread per chain (G = 1, register front), free-running: for ncu captures of fused_run_kernel<double, 1, ...>."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab
from aehmc_b200 import _engine
Cn = 65536
which = sys.argv[1] if len(sys.argv) > 1 else "schools"
model, eps = (ab.models.EightSchools(), 0.39) if which == "schools" else (ab.models.NealFunnel(10), 0.1)
state = ab.nuts.new_state(np.random.default_rng(0).standard_normal((Cn, 10)), model)
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    info, ex = _engine.run("nuts", model, np.ones(10), ab.RandomStream(seed=11), state, eps, max_ticks=300, return_counters=True, group=1)
    e1.record(); torch.cuda.synchronize()
leap = int(ex["counters"][0])
print(which, "ok leapfrogs", leap, "ms", e0.elapsed_time(e1), "evals/s %.3e" % (leap / e0.elapsed_time(e1) * 1e3))
