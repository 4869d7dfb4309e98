"""Trajectory builders (reference trajectory.py).

``static_integration`` is the HMC inner loop; the dynamic (NUTS) builders run inside the persistent
tick engine (csrc/engine.cuh) and are reached through ``nuts.new_kernel``; ``Diagnostics`` is the
per-transition result structure.
"""
from __future__ import annotations

from typing import NamedTuple

import torch

from .integrators import IntegratorState


class Diagnostics(NamedTuple):           # reference trajectory.py:379-384
    state: IntegratorState
    acceptance_probability: torch.Tensor  # [C] float64
    num_doublings: torch.Tensor           # [C] int32 (None for HMC)
    is_turning: torch.Tensor              # [C] bool  (None for HMC)
    is_diverging: torch.Tensor            # [C] bool


def static_integration(integrator, num_integration_steps):
    """reference trajectory.py:31-107: ``integrate(init_state, step_size) -> (state, updates)``.
    The whole fixed-length trajectory is one kernel launch."""

    def integrate(init_state, step_size):
        return integrator(init_state, step_size, n_steps=int(num_integration_steps)), {}

    return integrate
