"""HMC and NUTS transition kernels (test infrastructure).

Restates reference hmc.py:16-206 and nuts.py:14-155.  ``logprob_fn`` is a model
object from ``oracle.models`` (``potential_and_grad``); ``srng`` is a draws
provider from ``oracle.streams``.
"""
from __future__ import annotations

import math

import numpy as np

from . import hamiltonian, tree
from .hamiltonian import IntegratorState
from .tree import Diagnostics, ProposalState


def new_state(q, logprob_fn):
    """reference hmc.py:16-40 (nuts.new_state is the same function, nuts.py:14)."""
    U, g = logprob_fn.potential_and_grad(q)
    return IntegratorState(q, None, U, g)


def hmc_proposal(integrator, kinetic_energy, num_integration_steps, divergence_threshold):
    """reference hmc.py:129-206."""
    integrate = tree.static_integration(integrator, num_integration_steps)

    def propose(srng, state, step_size):
        new, _ = integrate(state, step_size)
        new = new._replace(momentum=-1.0 * new.momentum)                       # :185
        energy = state.potential_energy + kinetic_energy(state.momentum)
        new_energy = new.potential_energy + kinetic_energy(new.momentum)
        delta_energy = float(energy - new_energy)
        if math.isnan(delta_energy):
            delta_energy = -math.inf
        is_transition_divergent = abs(delta_energy) > divergence_threshold
        e = math.exp(delta_energy) if delta_energy < 700.0 else math.inf
        p_accept = min(max(e, 0.0), 1.0)                                       # :193
        do_accept = srng.hmc_accept(p_accept)
        final_state = new if do_accept else state                              # :195
        return Diagnostics(final_state, p_accept, None, None, bool(is_transition_divergent)), {}

    return propose


def hmc_new_kernel(srng, logprob_fn, divergence_threshold=1000):
    """reference hmc.py:43-126."""
    potential_fn = logprob_fn.potential_and_grad

    def step(state, step_size, inverse_mass_matrix, num_integration_steps):
        srng.begin_transition()
        momentum_generator, kinetic_energy_fn, _ = hamiltonian.gaussian_metric(inverse_mass_matrix)
        integrator = hamiltonian.velocity_verlet(potential_fn, kinetic_energy_fn)
        propose = hmc_proposal(integrator, kinetic_energy_fn, num_integration_steps, divergence_threshold)
        updated_state = state._replace(momentum=momentum_generator(srng))     # :122
        return propose(srng, updated_state, step_size)

    return step


def nuts_new_kernel(srng, logprob_fn, max_num_expansions=10, divergence_threshold=1000, exact_doubling=False):
    """reference nuts.py:17-155.  ``step`` returns (Diagnostics, extras)."""
    potential_fn = logprob_fn.potential_and_grad

    def step(state, step_size, inverse_mass_matrix):
        srng.begin_transition()
        momentum_generator, kinetic_energy_fn, uturn_check_fn = hamiltonian.gaussian_metric(
            inverse_mass_matrix
        )
        integrator = hamiltonian.velocity_verlet(potential_fn, kinetic_energy_fn)
        new_termination_state, update_termination_state, is_criterion_met = tree.iterative_uturn(
            uturn_check_fn
        )
        trajectory_integrator = tree.dynamic_integration(
            srng, integrator, kinetic_energy_fn, update_termination_state, is_criterion_met,
            divergence_threshold,
        )
        expand = tree.multiplicative_expansion(
            srng, trajectory_integrator, uturn_check_fn, max_num_expansions, exact_doubling
        )

        initial_state = state._replace(momentum=momentum_generator(srng))     # :113
        initial_termination_state = new_termination_state(initial_state.position, max_num_expansions)
        initial_energy = initial_state.potential_energy + kinetic_energy_fn(initial_state.momentum)
        initial_proposal = ProposalState(initial_state, initial_energy, 0.0, -np.inf)
        return expand(
            initial_proposal, initial_state, initial_state, initial_state.momentum,
            initial_termination_state, initial_energy, step_size,
        )

    return step
