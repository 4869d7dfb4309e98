"""Many transitions per call (the reference drives ``kernel`` from an outer ``aesara.scan``,
e.g. reference tests/test_hmc.py:296-324; here the loop stays on the device)."""
from __future__ import annotations

from . import _engine


def sample(kernel, state, step_size, inverse_mass_matrix, num_samples, *, num_integration_steps=None,
           store_draws=True, group=0):
    """Run ``num_samples`` transitions of every chain.  Returns (Diagnostics of the last transition,
    draws [num_samples, C, d] or None, stats [num_samples, C, 4], extras)."""
    spec = kernel.spec
    kw = dict(n_transitions=num_samples, divergence_threshold=spec["divergence_threshold"],
              store_draws=num_samples if store_draws else 0, group=group)
    if spec["kind"] == "nuts":
        kw["max_num_expansions"] = spec["max_num_expansions"]
    else:
        kw["num_integration_steps"] = int(num_integration_steps)
    info, extras = _engine.run(spec["kind"], spec["model"], inverse_mass_matrix, spec["srng"], state, step_size, **kw)
    return info, extras["draws"], extras["draw_stats"], extras
