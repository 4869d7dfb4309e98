"""c3 (logistic N=100k, D=128, 4096 chains) on the tensor-core path: a few ticks, for ncu launch lists."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab
from aehmc_b200 import _engine
N, D, Cn = 100000, 128, 4096
rng = np.random.default_rng(4)
X = torch.tensor(rng.standard_normal((N, D)), dtype=torch.float32).bfloat16().double().numpy()
beta = rng.standard_normal(D) / np.sqrt(D)
y = (rng.random(N) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
q0 = 0.1 * np.random.default_rng(6).standard_normal((Cn, D))
dt = torch.float32 if "--f32" in sys.argv else torch.float64
model = ab.models.LogisticRegression(X, y, 1.0, dtype=dt, tensor_core=True)
state = ab.nuts.new_state(q0, model)
for _ in range(2):
    info, ex = _engine.run("nuts", model, np.full(D, 4.0 / N), ab.RandomStream(seed=3), state, 0.4, max_ticks=3, return_counters=True)
torch.cuda.synchronize()
print("ok", int(ex["counters"][0]))
