"""dual_averaging_adaptation (reference step_size.py:9-100)."""
from __future__ import annotations

import torch

from . import algorithms


def dual_averaging_adaptation(target_acceptance_rate=0.8, gamma=0.05, t0=10, kappa=0.75):
    da_init, da_update = algorithms.dual_averaging(gamma, t0, kappa)

    def update(acceptance_probability, state):
        gradient = target_acceptance_rate - torch.as_tensor(acceptance_probability, dtype=torch.float64,
                                                            device=state.iterates.device)   # step_size.py:97
        return da_update(gradient, state)

    return da_init, update
