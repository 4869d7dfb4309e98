"""Proposals, iterative U-turn termination and trajectory builders (test infrastructure).

Restates reference proposals.py:11-174, termination.py:12-235 and
trajectory.py:31-735 as plain loops.  Every ``aesara.scan(..., until)`` becomes
a Python loop that runs the body at least once and stops after the iteration
whose condition is true, which is what Scan does.  The quirks listed in
SURVEY.md section 3.5 (Q1-Q12) are kept on purpose.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import numpy as np

from .hamiltonian import IntegratorState

NEG_INF = -math.inf


def logaddexp(a, b):
    """log(exp(a)+exp(b)) as the stabilised graph computes it: max + log(sum exp(x-max)),
    with exp(max) substituted when max is infinite (so -inf,-inf -> -inf)."""
    m = a if a > b else b
    if math.isnan(a) or math.isnan(b):
        return math.nan
    if math.isinf(m):
        e = math.exp(m) if m < 0 else math.inf
        s = e + e
        return m + (math.log(s) if s > 0 else NEG_INF)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def expit(x):
    if math.isnan(x):
        return math.nan
    if x < -709.0:          # exp(-x) would overflow: 1/(1+inf) = 0
        return 0.0
    return 1.0 / (1.0 + math.exp(-x))


# --------------------------------------------------------------------------
# proposals.py
# --------------------------------------------------------------------------
class ProposalState(NamedTuple):  # reference proposals.py:11-15
    state: IntegratorState
    energy: float
    weight: float
    sum_log_p_accept: float


def proposal_generator(kinetic_energy, divergence_threshold):
    """reference proposals.py:18-64."""

    def update(initial_energy, state):
        new_energy = state.potential_energy + kinetic_energy(state.momentum)
        delta_energy = float(initial_energy - new_energy)
        if math.isnan(delta_energy):
            delta_energy = NEG_INF
        is_transition_divergent = abs(delta_energy) > divergence_threshold
        weight = delta_energy
        log_p_accept = 0.0 if delta_energy > 0 else delta_energy
        return ProposalState(state, new_energy, weight, log_p_accept), bool(is_transition_divergent)

    return update


def maybe_update_proposal(do_accept, proposal, new_proposal):
    """reference proposals.py:137-174."""
    chosen = new_proposal if do_accept else proposal
    return ProposalState(
        state=chosen.state,
        energy=chosen.energy,
        weight=logaddexp(proposal.weight, new_proposal.weight),
        sum_log_p_accept=logaddexp(proposal.sum_log_p_accept, new_proposal.sum_log_p_accept),
    )


def progressive_uniform_sampling(draws, proposal, new_proposal, expansion=0, step=1):
    """reference proposals.py:72-102."""
    p_accept = expit(new_proposal.weight - proposal.weight)
    if math.isnan(p_accept):
        p_accept = 0.0
    do_accept = draws.uniform_accept(expansion, step, p_accept)
    return maybe_update_proposal(do_accept, proposal, new_proposal)


def progressive_biased_sampling(draws, proposal, new_proposal, expansion=0):
    """reference proposals.py:105-134."""
    diff = new_proposal.weight - proposal.weight
    e = math.exp(diff) if diff < 700.0 else math.inf
    p_accept = min(max(e, 0.0), 1.0) if not math.isnan(e) else math.nan
    do_accept = draws.biased_accept(expansion, p_accept)
    return maybe_update_proposal(do_accept, proposal, new_proposal)


def where_proposal(do_pick_left, left_proposal, right_proposal):
    """reference trajectory.py:717-735."""
    return left_proposal if do_pick_left else right_proposal


# --------------------------------------------------------------------------
# termination.py
# --------------------------------------------------------------------------
class TerminationState(NamedTuple):  # reference termination.py:12-16
    momentum_checkpoints: np.ndarray
    momentum_sum_checkpoints: np.ndarray
    min_index: int
    max_index: int


def _find_storage_indices(step):
    """reference termination.py:192-235, the two scans written out."""
    nc0, nc1 = int(step), -1
    for _ in range(int(step) + 1):            # count_subtrees
        stop = (nc0 & 1) == 0
        nc0, nc1 = nc0 // 2, nc1 + 1
        if stop:
            break
    num_subtrees = nc1
    nc0, nc1 = int(step) // 2, 0
    for _ in range(int(step) + 1):            # find_idx_max
        stop = nc0 == 0
        nc0, nc1 = nc0 // 2, nc1 + (nc0 & 1)
        if stop:
            break
    idx_max = nc1
    idx_min = idx_max - num_subtrees + 1
    return idx_min, idx_max


def iterative_uturn(is_turning_fn):
    """reference termination.py:19-189."""

    def new_state(position, max_num_doublings):                      # :43-83
        position = np.asarray(position)
        shape = (max_num_doublings,) if position.ndim == 0 else (max_num_doublings, position.shape[0])
        return TerminationState(np.zeros(shape), np.zeros(shape), 0, 0)

    def update(state, momentum_sum, momentum, step):                 # :85-131
        if step == 0:
            idx_min, idx_max = state.min_index, state.max_index      # stale indices (Q2)
        else:
            idx_min, idx_max = _find_storage_indices(step)
        if step % 2 == 0:
            mck = state.momentum_checkpoints.copy()
            sck = state.momentum_sum_checkpoints.copy()
            mck[idx_max] = momentum
            sck[idx_max] = momentum_sum
        else:
            mck, sck = state.momentum_checkpoints, state.momentum_sum_checkpoints
        return TerminationState(mck, sck, idx_min, idx_max)

    def is_iterative_turning(state, momentum_sum, momentum):         # :133-187
        i = state.max_index
        criterion = False
        for _ in range(state.max_index + 2):
            subtree_momentum_sum = (
                momentum_sum - state.momentum_sum_checkpoints[i] + state.momentum_checkpoints[i]
            )
            criterion = is_turning_fn(state.momentum_checkpoints[i], momentum, subtree_momentum_sum)
            reached_max_iteration = (i - 1) < state.min_index
            i -= 1
            if criterion or reached_max_iteration:
                break
        if state.max_index < state.min_index:
            return False
        return bool(criterion)

    return new_state, update, is_iterative_turning


# --------------------------------------------------------------------------
# trajectory.py
# --------------------------------------------------------------------------
def static_integration(integrator, num_integration_steps):
    """reference trajectory.py:31-107."""

    def integrate(init_state, step_size):
        state = init_state
        for _ in range(int(num_integration_steps)):
            state = integrator(state, step_size)
        return state, {}

    return integrate


class SubtreeResult(NamedTuple):
    proposal: ProposalState
    state: IntegratorState
    momentum_sum: object
    termination_state: TerminationState
    trajectory_length: int
    is_diverging: bool
    has_terminated: bool


def dynamic_integration(
    draws, integrator, kinetic_energy, update_termination_state, is_criterion_met, divergence_threshold
):
    """reference trajectory.py:119-376."""
    generate_proposal = proposal_generator(kinetic_energy, divergence_threshold)

    def integrate(previous_last_state, direction, termination_state, max_num_steps, step_size,
                  initial_energy, expansion=0):
        # one step away to start the sub-trajectory (:276-284)
        state = integrator(previous_last_state, direction * step_size)
        proposal, is_diverging = generate_proposal(initial_energy, state)
        momentum_sum = state.momentum
        termination_state = update_termination_state(termination_state, momentum_sum, state.momentum, 0)
        full_initial = SubtreeResult(proposal, state, momentum_sum, termination_state, 1, is_diverging, False)
        first_diverged = is_diverging

        # the scan over arange(1, 1 + max_num_steps) (:307-332); it executes even when the
        # first step diverged (its RNG update is a graph output), result selected after (:336)
        last_state, length = state, 1
        has_terminated = False
        for step in range(1, 1 + int(max_num_steps)):
            new_state = integrator(last_state, direction * step_size)
            new_proposal, is_diverging = generate_proposal(initial_energy, new_state)
            proposal = progressive_uniform_sampling(draws, proposal, new_proposal, expansion, step)
            momentum_sum = momentum_sum + new_state.momentum
            termination_state = update_termination_state(
                termination_state, momentum_sum, new_state.momentum, step
            )
            has_terminated = is_criterion_met(termination_state, momentum_sum, new_state.momentum)
            last_state, length = new_state, length + 1
            if is_diverging or has_terminated:
                break
        full_last = SubtreeResult(
            proposal, last_state, momentum_sum, termination_state, length, is_diverging, has_terminated
        )
        return (full_initial if first_diverged else full_last), {}

    return integrate


class Diagnostics(NamedTuple):  # reference trajectory.py:379-384
    state: IntegratorState
    acceptance_probability: object
    num_doublings: object
    is_turning: object
    is_diverging: object


class ExpansionTrace(NamedTuple):
    """One executed expansion (the rows of SURVEY.md section 3.7's trace)."""
    expansion: int
    go_right: bool
    subtree_length: int
    is_diverging: bool
    subtree_terminated: bool
    is_turning: bool
    proposal_position: object


def multiplicative_expansion(draws, trajectory_integrator, uturn_check_fn, max_num_expansions, exact_doubling=False):
    """reference trajectory.py:396-714.  Returns the values of the LAST executed
    expansion (what nuts.py:138-151 extracts) plus a per-expansion trace.
    ``exact_doubling`` is NOT reference behaviour: it shortens every sub-tree from the reference's ``2**k + 1``
    leapfrogs to the balanced ``2**k`` (the checker of the engine option of the same name)."""
    sub_steps = (lambda step: 2 ** step - 1) if exact_doubling else (lambda step: 2 ** step)

    def expand(proposal, left_state, right_state, momentum_sum, termination_state,
               initial_energy, step_size):
        trace = []
        acceptance_probability, num_doublings = None, 0
        is_diverging = is_turning = False
        n_leapfrog = 0
        for step in range(int(max_num_expansions)):
            do_go_right = draws.direction(step)                                  # :516
            direction = 1.0 if do_go_right else -1.0
            start_state = right_state if do_go_right else left_state
            sub, _ = trajectory_integrator(
                start_state, direction, termination_state, sub_steps(step), step_size, initial_energy,
                expansion=step,
            )
            new_proposal, new_state = sub.proposal, sub.state
            termination_state = sub.termination_state
            is_diverging, has_subtree_terminated = sub.is_diverging, sub.has_terminated
            n_leapfrog += sub.trajectory_length

            if do_go_right:                                                      # :540-545
                right_state = new_state
            else:
                left_state = new_state
            momentum_sum = momentum_sum + sub.momentum_sum                       # :546
            acceptance_probability = (                                           # :551-553
                math.exp(new_proposal.sum_log_p_accept) / sub.trajectory_length
            )
            updated_proposal = proposal._replace(                                # :560-564
                sum_log_p_accept=logaddexp(new_proposal.sum_log_p_accept, proposal.sum_log_p_accept)
            )
            sampled = progressive_biased_sampling(draws, proposal, new_proposal, step)  # always drawn
            proposal = where_proposal(is_diverging or has_subtree_terminated, updated_proposal, sampled)
            is_turning = uturn_check_fn(left_state.momentum, right_state.momentum, momentum_sum)
            num_doublings = step + 1
            trace.append(ExpansionTrace(step, do_go_right, sub.trajectory_length, is_diverging,
                                        has_subtree_terminated, is_turning, proposal.state.position))
            if is_diverging or is_turning or has_subtree_terminated:             # :577
                break
        diagnostics = Diagnostics(proposal.state, acceptance_probability, num_doublings,
                                  is_turning, is_diverging)
        extras = {
            "trace": trace, "n_leapfrog": n_leapfrog, "proposal": proposal,
            "left_state": left_state, "right_state": right_state, "momentum_sum": momentum_sum,
            "termination_state": termination_state,
        }
        return diagnostics, extras

    return expand
