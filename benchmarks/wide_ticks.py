"""Fused persistent NUTS kernel on iid Gaussian d=128, 131072 chains (for ncu captures)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab
from aehmc_b200 import _engine
d, Cn = 128, 131072
rng = np.random.default_rng(0)
sigma = np.exp(0.3 * rng.standard_normal(d))
model = ab.models.IIDGaussian(np.zeros(d), sigma)
state = ab.nuts.new_state(rng.standard_normal((Cn, d)) * sigma, model)
for _ in range(2):
    info, ex = _engine.run("nuts", model, sigma ** 2, ab.RandomStream(seed=5), state, 1.2 / d ** 0.25, n_transitions=5, return_counters=True)
torch.cuda.synchronize()
print("ok", int(ex["counters"][0]))
