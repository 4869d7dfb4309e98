"""Integrator and metric restatement (test infrastructure).

Follows reference integrators.py:7-75 and metrics.py:10-106.  Arithmetic is
written in the association order of the reference expressions so that the
scalar-metric case is bit-identical to the compiled Aesara graph (checked by the
README golden).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import scipy.linalg

from . import streams


class IntegratorState(NamedTuple):  # reference integrators.py:7-11
    position: object
    momentum: object
    potential_energy: object
    potential_energy_grad: object


def new_integrator_state(potential_fn, position, momentum):
    """reference integrators.py:14-24; ``potential_fn(q) -> (U, g)``."""
    U, g = potential_fn(position)
    return IntegratorState(position, momentum, U, g)


def velocity_verlet(potential_fn, kinetic_energy_fn):
    """reference integrators.py:27-75.  ``kinetic_energy_fn.grad(p)`` stands in
    for ``aesara.grad(kinetic_energy_fn(p), p)`` (integrators.py:61)."""
    a1 = 0
    b1 = 0.5
    a2 = 1 - 2 * a1

    def one_step(state, step_size):
        half = b1 * step_size
        momentum = state.momentum - half * state.potential_energy_grad      # :59
        position = state.position + (a2 * step_size) * kinetic_energy_fn.grad(momentum)  # :61-62
        U, g = potential_fn(position)                                        # :64-65
        momentum = momentum - half * g                                       # :66
        return IntegratorState(position, momentum, U, g)

    return one_step


def gaussian_metric(inverse_mass_matrix):
    """reference metrics.py:44-106 -> (momentum_generator, kinetic_energy, is_turning)."""
    imm = np.asarray(inverse_mass_matrix, dtype=np.float64)
    if imm.ndim == 0:
        shape = ()
        mass_matrix_sqrt = np.sqrt(np.reciprocal(imm))
        matmul = lambda a, b: a * b
        dot = lambda a, b: a * b
        vel_grad = lambda p: imm * p
    elif imm.ndim == 1:
        shape = (imm.shape[0],)
        mass_matrix_sqrt = np.sqrt(np.reciprocal(imm))
        matmul = lambda a, b: a * b
        dot = np.dot
        vel_grad = lambda p: imm * p
    elif imm.ndim == 2:
        shape = (imm.shape[0],)
        L = np.linalg.cholesky(imm)                                          # :56
        mass_matrix_sqrt = scipy.linalg.solve_triangular(                    # :58 (lower, trans)
            L, np.eye(shape[0]), lower=True, trans="T"
        )
        matmul = np.dot
        dot = np.dot
        sym = 0.5 * (imm + imm.T)   # gradient of 0.5 p^T A p
        vel_grad = lambda p: sym @ p
    else:
        raise ValueError(
            f"Expected a mass matrix of dimension 1 (diagonal) or 2, got {imm.ndim}"
        )

    def momentum_generator(draws):                                           # :65-68
        z = draws.normal(shape)
        return matmul(mass_matrix_sqrt, z)

    def kinetic_energy(momentum):                                            # :70-73
        velocity = matmul(imm, momentum)
        return 0.5 * dot(velocity, momentum)

    kinetic_energy.grad = vel_grad

    def is_turning(momentum_left, momentum_right, momentum_sum):             # :75-104
        velocity_left = matmul(imm, momentum_left)
        velocity_right = matmul(imm, momentum_right)
        rho = momentum_sum - (momentum_right + momentum_left) / 2
        turning_at_left = dot(velocity_left, rho) <= 0
        turning_at_right = dot(velocity_right, rho) <= 0
        if streams.MARGIN_LOG is not None:
            nr = np.sqrt(np.sum(np.square(rho)))
            for v in (velocity_left, velocity_right) if nr > 0 else ():      # rho == 0: a one-state trajectory, no tie
                streams.MARGIN_LOG.append(("uturn", abs(float(dot(v, rho))) / (np.sqrt(np.sum(np.square(v))) * nr + 1e-300)))
        return bool(turning_at_left | turning_at_right)

    return momentum_generator, kinetic_energy, is_turning
