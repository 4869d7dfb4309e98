"""HMC kernel (reference hmc.py) for many chains at once."""
from __future__ import annotations

from . import _engine
from .integrators import IntegratorState


def new_state(q, logprob_fn):
    """reference hmc.py:16-40: state of every chain at positions q[C, d]."""
    U, g = logprob_fn.potential_and_grad(q)
    from . import backend
    return IntegratorState(backend.as_device(q, logprob_fn.dtype, logprob_fn.device), None, U, g)


def new_kernel(srng, logprob_fn, divergence_threshold=1000):
    """reference hmc.py:43-126.  ``step(state, step_size, inverse_mass_matrix, num_integration_steps)``
    -> (Diagnostics, updates).  Momentum draw, L fused leapfrogs and the Metropolis accept of every
    chain run in one persistent kernel."""

    def step(state, step_size, inverse_mass_matrix, num_integration_steps):
        info, extras = _engine.run("hmc", logprob_fn, inverse_mass_matrix, srng, state, step_size,
                                   divergence_threshold=divergence_threshold,
                                   num_integration_steps=int(num_integration_steps))
        return info, {"n_leapfrog": extras["n_leapfrog"]}

    step.spec = dict(kind="hmc", srng=srng, model=logprob_fn, divergence_threshold=divergence_threshold)
    return step
