"""Generates tests/golden/nuts_moments.json: long-run moments of the ORACLE's NUTS chain (i.e. of the
reference algorithm as published, including its 2**k + 1 sub-tree length) on the
reference's own MCSE target (tests/test_hmc.py:170-187).  The reference algorithm is not exactly
invariant -- see DESIGN.md "A property of the reference" -- so the native-RNG GPU test compares with
these moments, not with the analytic posterior.  Run from the repo root:  python tests/golden/make_nuts_moments.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import kernels, models, streams  # noqa: E402

loc = np.array([0.0, 3.0]); scale = np.array([1.0, 2.0]); rho = 0.5
cov = np.array([[scale[0] ** 2, rho * scale[0] * scale[1]], [rho * scale[0] * scale[1], scale[1] ** 2]])
model = models.CorrelatedGaussian(loc, np.linalg.inv(cov))
out = {"target": {"loc": loc.tolist(), "scale": scale.tolist(), "rho": rho}, "cases": []}
for eps, imm in ((1.0, [1.0, 1.0]), (0.3, [1.0, 4.0])):
    xs = []
    for seed in range(4):
        srng = streams.StreamDraws(seed, "nuts")
        kernel = kernels.nuts_new_kernel(srng, model)
        state = kernels.new_state(np.array([1.0, 1.0]), model)
        for i in range(10500):
            info, _ = kernel(state, eps, np.array(imm))
            state = info.state._replace(momentum=None)
            if i >= 500:
                xs.append(info.state.position)
    xs = np.array(xs)
    out["cases"].append({"step_size": eps, "inverse_mass_matrix": imm, "n": len(xs), "mean": xs.mean(0).tolist(),
                         "var": xs.var(0).tolist(), "corr": float(np.corrcoef(xs.T)[0, 1])})
    print(out["cases"][-1], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nuts_moments.json"), "w"), indent=1)
