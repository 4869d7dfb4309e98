"""Why does the rank-normalised R-hat of the README example differ from the plain one?  Prints the R-hat variants and
the spread of the per-chain scales (per-chain adaptation + the reference's 2**k + 1 sub-tree length)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aehmc_b200 as ab  # noqa: E402

model = ab.models.IIDGaussian([0.0], [1.0])
for label, kw, pooled in (("reference sub-trees, per-chain warm-up", {}, False), ("reference sub-trees, pooled warm-up", {}, True),
                          ("exact_doubling, per-chain warm-up", {"exact_doubling": True}, False)):
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=0), model, **kw)
    state = ab.nuts.new_state(np.zeros((2048, 1)), model)
    state, (eps, imm), _ = ab.window_adaptation.run(kernel, state, num_steps=200, pooled=pooled)
    metric = imm if pooled else ab.metrics.per_chain(imm)
    info, draws, stats, _ = ab.sampling.sample(kernel, state, eps, metric, 400, thin=1)
    x = draws[:, :, 0].double()
    sd = x.std(0)
    d = ab.diagnostics
    print(label)
    print("  eps quantiles", np.quantile(eps.cpu().numpy(), [0.01, 0.5, 0.99]), "imm quantiles",
          np.quantile(imm.double().cpu().numpy(), [0.01, 0.5, 0.99]))
    print("  per-chain sd quantiles", np.quantile(sd.cpu().numpy(), [0.01, 0.25, 0.5, 0.75, 0.99]), "pooled sd", float(x.std()))
    print("  rhat identity", d.rhat(draws, method="identity"), "split", d.rhat(draws, method="split"), "rank", d.rhat(draws))
    print("  ess bulk", d.ess(draws), "ess mean", d.ess(draws, method="mean"), "of", draws.shape[0] * draws.shape[1])
    thin = draws[::10]
    print("  thinned x10: rhat rank", d.rhat(thin), "identity", d.rhat(thin, method="identity"), "ess bulk", d.ess(thin))
