"""Warm-up adaptation restatement (test infrastructure).

Restates reference algorithms.py:17-204, step_size.py:9-100,
mass_matrix.py:12-120 and window_adaptation.py:17-327, including the quirks
Q16-Q19 of SURVEY.md section 3.5 (the shrinkage point is the step size itself,
the averaged iterate uses the OLD iterate, the first step size is exp(0)).
"""
from __future__ import annotations

import math
from typing import NamedTuple

import numpy as np

from .hamiltonian import IntegratorState


class DualAveragingState(NamedTuple):  # reference algorithms.py:9-14
    step: int
    iterates: float
    iterates_avg: float
    gradient_avg: float
    shrinkage_pts: float


def dual_averaging(gamma=0.05, t0=10, kappa=0.75):
    """reference algorithms.py:17-117."""

    def init(mu):
        return DualAveragingState(1, 0.0, 0.0, 0.0, mu)

    def update(gradient, state):
        eta = 1.0 / (state.step + t0)
        new_gradient_avg = (1.0 - eta) * state.gradient_avg + eta * gradient
        new_x = state.shrinkage_pts - (math.sqrt(state.step) / gamma) * new_gradient_avg
        x_eta = float(state.step) ** (-kappa)
        new_x_avg = x_eta * state.iterates + (1.0 - x_eta) * state.iterates_avg
        return state._replace(
            step=state.step + 1, iterates=new_x, iterates_avg=new_x_avg, gradient_avg=new_gradient_avg
        )

    return init, update


def welford_covariance(compute_covariance):
    """reference algorithms.py:120-204."""

    def init(n_dims):
        if n_dims == 0:
            return 0.0, 0.0, 0
        mean = np.zeros((n_dims,))
        m2 = np.zeros((n_dims, n_dims)) if compute_covariance else np.zeros((n_dims,))
        return mean, m2, 0

    def update(value, mean, m2, sample_size):
        sample_size = sample_size + 1
        delta = value - mean
        mean = mean + delta / sample_size
        updated_delta = value - mean
        if compute_covariance and np.ndim(mean) > 0:
            m2 = m2 + np.outer(updated_delta, delta)
        else:
            m2 = m2 + updated_delta * delta
        return mean, m2, sample_size

    def final(m2, sample_size):
        with np.errstate(divide="ignore", invalid="ignore"):
            return m2 / np.float64(sample_size - 1)

    return init, update, final


def dual_averaging_adaptation(target_acceptance_rate=0.8, gamma=0.05, t0=10, kappa=0.75):
    """reference step_size.py:9-100."""
    da_init, da_update = dual_averaging(gamma, t0, kappa)

    def update(acceptance_probability, state):
        return da_update(target_acceptance_rate - acceptance_probability, state)

    return da_init, update


def covariance_adaptation(is_mass_matrix_full=False):
    """reference mass_matrix.py:12-120."""
    wc_init, wc_update, wc_final = welford_covariance(is_mass_matrix_full)

    def init(n_dims):
        if n_dims == 0:
            imm = 1.0
        elif is_mass_matrix_full:
            imm = np.eye(n_dims)
        else:
            imm = np.ones((n_dims,))
        return imm, wc_init(n_dims)

    def update(position, wc_state):
        return wc_update(position, *wc_state)

    def final(wc_state):
        _, m2, n = wc_state
        covariance = wc_final(m2, n)
        scaled = (n / (n + 5)) * covariance
        shrinkage = 1e-3 * (5 / (n + 5))
        if np.ndim(covariance) > 0 and is_mass_matrix_full:
            return scaled + shrinkage * np.eye(covariance.shape[0])
        return scaled + shrinkage

    return init, update, final


def build_schedule(num_steps, initial_buffer_size=75, final_buffer_size=50, first_window_size=25):
    """Stan's three-stage warm-up schedule as the reference lays it out
    (window_adaptation.py:230-327): a list of (stage, is_middle_window_end)."""
    if num_steps < 20:
        return [(0, False) for _ in range(num_steps)]
    if initial_buffer_size + first_window_size + final_buffer_size > num_steps:
        initial_buffer_size = int(0.15 * num_steps)
        final_buffer_size = int(0.1 * num_steps)
        first_window_size = num_steps - initial_buffer_size - final_buffer_size
    slow_end = num_steps - final_buffer_size
    stage = np.zeros(num_steps, dtype=np.int64)
    window_end = np.zeros(num_steps, dtype=bool)
    stage[initial_buffer_size:slow_end] = 1
    start, size = initial_buffer_size, first_window_size
    while start < slow_end:
        this_size = size
        if 3 * size <= slow_end - start:
            size = 2 * size
        else:
            this_size = slow_end - start
        start += this_size
        window_end[start - 1] = True
    return [(int(s), bool(e)) for s, e in zip(stage, window_end)]


def window_adaptation(num_steps, is_mass_matrix_full=False, initial_step_size=1.0,
                      target_acceptance_rate=0.80):
    """reference window_adaptation.py:119-227."""
    mm_init, mm_update, mm_final = covariance_adaptation(is_mass_matrix_full)
    da_init, da_update = dual_averaging_adaptation(target_acceptance_rate)
    schedule = build_schedule(num_steps)

    def init(initial_chain_state):
        pos = np.asarray(initial_chain_state.position)
        num_dims = 0 if pos.ndim == 0 else pos.shape[0]
        imm, mm_state = mm_init(num_dims)
        da_state = da_init(initial_step_size)          # mu = the step size itself (Q16)
        step_size = math.exp(da_state.iterates)
        return (da_state, mm_state), (step_size, imm)

    def fast_update(p_accept, warmup_state, parameters):
        da_state, mm_state = warmup_state
        new_da = da_update(p_accept, da_state)
        return (new_da, mm_state), (math.exp(new_da.iterates), parameters[1])

    def slow_update(position, p_accept, warmup_state, parameters):
        da_state, mm_state = warmup_state
        new_da = da_update(p_accept, da_state)
        new_mm = mm_update(position, mm_state)
        return (new_da, new_mm), (math.exp(new_da.iterates), parameters[1])

    def slow_final(warmup_state):
        da_state, mm_state = warmup_state
        imm = mm_final(mm_state)
        num_dims = 0 if np.ndim(imm) == 0 else np.shape(imm)[0]
        _, new_mm_state = mm_init(num_dims)
        step_size = math.exp(da_state.iterates)
        return (da_init(step_size), new_mm_state), (step_size, imm)

    def update(step, warmup_state, parameters, chain_info):
        stage, is_middle_window_end = schedule[step]
        if stage == 0:
            warmup_state, parameters = fast_update(
                chain_info.acceptance_probability, warmup_state, parameters)
        else:
            warmup_state, parameters = slow_update(
                chain_info.state.position, chain_info.acceptance_probability, warmup_state, parameters)
        if is_middle_window_end:
            warmup_state, parameters = slow_final(warmup_state)
        if step == num_steps - 1:
            parameters = (math.exp(warmup_state[0].iterates_avg), parameters[1])
        return warmup_state, parameters

    return init, update


def run(kernel, initial_state, num_steps=1000, *, is_mass_matrix_full=False,
        initial_step_size=1.0, target_acceptance_rate=0.80, trace=None):
    """reference window_adaptation.py:17-116.  ``kernel(state, step_size, imm)``
    returns (Diagnostics, extras)."""
    init_adapt, update_adapt = window_adaptation(
        num_steps, is_mass_matrix_full, initial_step_size, target_acceptance_rate)
    warmup_state, parameters = init_adapt(initial_state)
    chain_state = initial_state
    for warmup_step in range(num_steps):
        chain_info, extras = kernel(chain_state, *parameters)
        warmup_state, parameters = update_adapt(warmup_step, warmup_state, parameters, chain_info)
        chain_state = IntegratorState(
            chain_info.state.position, None, chain_info.state.potential_energy,
            chain_info.state.potential_energy_grad)
        if trace is not None:
            trace.append((chain_info, extras, parameters))
    return chain_state, parameters, {}


def merge_welford(n_a, mean_a, m2_a, n_b, mean_b, m2_b):
    """Chan's pairwise merge of two Welford states: what folding the second group's values one by one with
    welford_covariance.update (reference algorithms.py:187-197) converges to, in one step."""
    if n_b == 0:
        return n_a, mean_a, m2_a
    if n_a == 0:
        return n_b, np.array(mean_b, dtype=np.float64), np.array(m2_b, dtype=np.float64)
    n = n_a + n_b
    delta = mean_b - mean_a
    cross = np.outer(delta, delta) if np.ndim(m2_a) == 2 else delta * delta
    return n, mean_a + delta * (n_b / n), m2_a + m2_b + cross * (n_a * n_b / n)


def run_pooled(kernels, initial_states, num_steps, *, is_mass_matrix_full=False, initial_step_size=1.0,
               target_acceptance_rate=0.80, trace=None):
    """Pooled warm-up of many chains of one target (beyond the reference, which adapts one chain at a time): per-chain
    dual averaging exactly as window_adaptation.update (reference window_adaptation.py:194-227), ONE inverse mass matrix
    re-estimated at each slow-window end from the slow-stage positions of all chains with the reference's own
    welford_covariance / covariance_adaptation.final (values folded one by one: the defining recurrence)."""
    mm_init, mm_update, mm_final = covariance_adaptation(is_mass_matrix_full)
    da_init, da_update = dual_averaging_adaptation(target_acceptance_rate)
    schedule = build_schedule(num_steps)
    n_chains = len(kernels)
    d = np.asarray(initial_states[0].position).shape[0]
    imm, mm_state = mm_init(d)
    da_states = [da_init(initial_step_size) for _ in range(n_chains)]
    step_sizes = [math.exp(s.iterates) for s in da_states]
    states = list(initial_states)
    for step in range(num_steps):
        stage, is_end = schedule[step]
        infos = []
        for c in range(n_chains):
            info, extras = kernels[c](states[c], step_sizes[c], imm)
            infos.append((info, extras))
            da_states[c] = da_update(info.acceptance_probability, da_states[c])
            step_sizes[c] = math.exp(da_states[c].iterates)
            states[c] = IntegratorState(info.state.position, None, info.state.potential_energy,
                                        info.state.potential_energy_grad)
        if stage == 1:
            for c in range(n_chains):
                mm_state = mm_update(infos[c][0].state.position, mm_state)
        if is_end:
            imm = mm_final(mm_state)
            _, mm_state = mm_init(d)
            da_states = [da_init(step_sizes[c]) for c in range(n_chains)]
        if step == num_steps - 1:
            step_sizes = [math.exp(s.iterates_avg) for s in da_states]
        if trace is not None:
            trace.append((infos, list(step_sizes), imm))
    return states, (np.array(step_sizes), imm), {}
