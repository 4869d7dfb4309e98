"""Iterative NUTS kernel (reference nuts.py) for many chains at once."""
from __future__ import annotations

from . import _engine, hmc

new_state = hmc.new_state          # reference nuts.py:14


def new_kernel(srng, logprob_fn, max_num_expansions=10, divergence_threshold=1000, exact_doubling=False):
    """reference nuts.py:17-155.  ``step(state, step_size, inverse_mass_matrix)`` -> (Diagnostics, updates).
    The multiplicative expansion, iterative U-turn checkpoints, progressive sampling and divergence
    checks of every chain run in the tick engine (csrc/engine.cuh); ``updates`` carries ``n_leapfrog``.

    ``exact_doubling`` (not in the reference, default off): the reference builds sub-trees of ``2**k + 1``
    leapfrogs (trajectory.py:276,302,307), which this build reproduces decision for decision but which makes the
    sampler slightly biased; ``exact_doubling=True`` uses balanced sub-trees of ``2**k`` leapfrogs,
    for which the iterative U-turn criterion and the biased progressive sampling leave the target invariant."""

    def step(state, step_size, inverse_mass_matrix):
        info, extras = _engine.run("nuts", logprob_fn, inverse_mass_matrix, srng, state, step_size,
                                   max_num_expansions=max_num_expansions,
                                   divergence_threshold=divergence_threshold, exact_doubling=exact_doubling)
        return info, {"n_leapfrog": extras["n_leapfrog"]}

    step.spec = dict(kind="nuts", srng=srng, model=logprob_fn, max_num_expansions=max_num_expansions,
                     divergence_threshold=divergence_threshold, exact_doubling=bool(exact_doubling))
    return step
