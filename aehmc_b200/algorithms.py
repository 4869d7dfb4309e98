"""dual_averaging / welford_covariance (reference algorithms.py) over batched chains."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch

from . import _lib, backend


class DualAveragingState(NamedTuple):     # reference algorithms.py:9-14
    step: torch.Tensor          # [C] int64
    iterates: torch.Tensor      # [C] float64
    iterates_avg: torch.Tensor
    gradient_avg: torch.Tensor
    shrinkage_pts: torch.Tensor


def dual_averaging(gamma=0.05, t0=10, kappa=0.75):
    """reference algorithms.py:17-117: ``init(mu[C]) -> state`` and ``update(gradient[C], state)``."""
    lib = _lib.load()

    def init(mu):
        mu = torch.as_tensor(mu, dtype=torch.float64, device=backend.device(getattr(mu, "device", None) if isinstance(mu, torch.Tensor) and mu.is_cuda else None))
        mu = mu.reshape(-1).contiguous()
        z = torch.zeros_like(mu)
        return DualAveragingState(torch.ones(mu.shape[0], dtype=torch.int64, device=mu.device), z, z.clone(),
                                  z.clone(), mu)

    def update(gradient, state):
        dev = state.iterates.device
        Cn = state.iterates.shape[0]
        # the kernel takes p_accept and a target: gradient = target - p_accept with target 0
        neg = (-backend.as_device(gradient, torch.float64, dev)).contiguous()
        step, x, xa, ga = (t.clone() for t in (state.step, state.iterates, state.iterates_avg, state.gradient_avg))
        _lib.check(lib.b2h_dual_averaging_update(backend.context(dev), backend.ptr(neg), C.c_double(0.0),
                                                 C.c_double(gamma), C.c_double(t0), C.c_double(kappa),
                                                 backend.ptr(step), backend.ptr(x), backend.ptr(xa), backend.ptr(ga),
                                                 backend.ptr(state.shrinkage_pts), C.c_int64(Cn)))
        return DualAveragingState(step, x, xa, ga, state.shrinkage_pts)

    return init, update


def welford_covariance(compute_covariance):
    """reference algorithms.py:120-204: ``init(n_dims, num_chains) -> (mean, m2, n)``, ``update``, ``final``."""
    lib = _lib.load()
    full = 1 if compute_covariance else 0

    def init(n_dims, num_chains=1, dtype=torch.float64, device=None):
        dev = backend.device(device)
        mean = torch.zeros((num_chains, n_dims), dtype=dtype, device=dev)
        m2 = torch.zeros((num_chains, n_dims, n_dims) if compute_covariance else (num_chains, n_dims),
                         dtype=dtype, device=dev)
        return mean, m2, torch.zeros(num_chains, dtype=torch.int64, device=dev)

    def update(value, mean, m2, sample_size):
        dev, dt = mean.device, mean.dtype
        value = backend.as_device(value, dt, dev)
        mean, m2, n = mean.clone(), m2.clone(), sample_size.clone()
        Cn, d = mean.shape
        _lib.check(lib.b2h_welford_update(backend.context(dev), backend.code(dt), backend.ptr(value), backend.ptr(mean),
                                          backend.ptr(m2), backend.ptr(n), C.c_int64(Cn), C.c_int64(d), C.c_int32(full)))
        return mean, m2, n

    def final(m2, sample_size):
        n = (sample_size - 1).to(m2.dtype)
        return m2 / n.reshape((-1,) + (1,) * (m2.ndim - 1))

    return init, update, final
