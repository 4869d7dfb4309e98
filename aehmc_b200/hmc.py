"""HMC kernel (reference hmc.py) for many chains at once."""
from __future__ import annotations

import ctypes as C

import torch

from . import _engine, _lib, backend, trajectory
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


def hmc_proposal(integrator, kinetic_energy, num_integration_steps, divergence_threshold):
    """reference hmc.py:129-206 -> ``propose(srng, state, step_size) -> (Diagnostics, updates)``.
    ``integrator`` comes from ``integrators.velocity_verlet`` and ``kinetic_energy`` from ``metrics.gaussian_metric``.
    The fixed-length trajectory is one launch (``trajectory.static_integration``); the momentum flip, the energy
    difference, the divergence flag and the Metropolis accept are one more (``b2h_hmc_accept``).  ``state`` must carry a
    momentum (``hmc.new_kernel`` draws it with ``momentum_generator`` first, hmc.py:121-122)."""
    integrate = trajectory.static_integration(integrator, num_integration_steps)
    lib = _lib.load()

    def propose(srng, state, step_size):
        if state.momentum is None:
            raise ValueError("hmc_proposal: the state needs a momentum (see metrics.gaussian_metric momentum_generator)")
        new_state, updates = integrate(state, step_size)
        dev, dt = new_state.position.device, new_state.position.dtype
        Cn, d = new_state.position.shape
        old = IntegratorState(*[backend.as_device(t, dt, dev) for t in state])
        K_old = kinetic_energy(old.momentum).to(dt).contiguous()
        K_new = kinetic_energy(new_state.momentum).to(dt).contiguous()      # K(-p) == K(p)
        u = srng.uniform(Cn, dev)
        p_accept = torch.empty(Cn, dtype=torch.float64, device=dev)
        div = torch.empty(Cn, dtype=torch.uint8, device=dev)
        st = lambda s: _lib.State(s.position.data_ptr(), s.momentum.data_ptr(), s.potential_energy_grad.data_ptr(),
                                  s.potential_energy.data_ptr())
        o, n = st(old), st(new_state)
        _lib.check(lib.b2h_hmc_accept(backend.context(dev), backend.code(dt), C.byref(o), C.byref(n), backend.ptr(K_old),
                                      backend.ptr(K_new), backend.ptr(u), C.c_double(float(divergence_threshold)),
                                      backend.ptr(p_accept), backend.ptr(div), C.c_int64(Cn), C.c_int64(d)))
        return trajectory.Diagnostics(new_state, p_accept, None, None, div.bool()), updates

    return propose
