"""Proposal bookkeeping (reference proposals.py) as batched device primitives.

Inside ``nuts.new_kernel`` these steps run fused in the tick engine (csrc/engine.cuh post_gradient);
the functions here expose the same closures for composition and testing.  Chains are the leading
axis; weights and sum_log_p_accept are float64 as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch

from . import _lib, backend
from .integrators import IntegratorState


class ProposalState(NamedTuple):          # reference proposals.py:11-15
    state: IntegratorState
    energy: torch.Tensor                  # [C] dtype
    weight: torch.Tensor                  # [C] float64
    sum_log_p_accept: torch.Tensor        # [C] float64


def proposal_generator(kinetic_energy, divergence_threshold):
    """reference proposals.py:18-64: ``update(initial_energy, state) -> (ProposalState, is_diverging)``."""
    lib = _lib.load()

    def update(initial_energy, state):
        K = kinetic_energy(state.momentum)
        U = state.potential_energy
        dev, dt = U.device, U.dtype
        Cn = U.shape[0]
        E0 = backend.as_device(initial_energy, dt, dev).expand(Cn).contiguous()
        energy = torch.empty_like(U)
        weight = torch.empty(Cn, dtype=torch.float64, device=dev)
        lpa = torch.empty(Cn, dtype=torch.float64, device=dev)
        div = torch.empty(Cn, dtype=torch.uint8, device=dev)
        _lib.check(lib.b2h_proposal_update(backend.context(dev), backend.code(dt), backend.ptr(E0), backend.ptr(U),
                                           backend.ptr(K.to(dt).contiguous()), C.c_double(float(divergence_threshold)),
                                           backend.ptr(energy), backend.ptr(weight), backend.ptr(lpa),
                                           backend.ptr(div), C.c_int64(Cn)))
        return ProposalState(state, energy, weight, lpa), div.bool()

    return update


def _select(mask, a, b):
    """row-wise where() on the device (b2h_select_rows)."""
    lib = _lib.load()
    a = a.contiguous()
    b = b.contiguous()
    out = torch.empty_like(a)
    Cn = a.shape[0]
    d = a.numel() // max(Cn, 1)
    if a.dtype == torch.float64 or a.dtype == torch.float32:
        _lib.check(lib.b2h_select_rows(backend.context(a.device), backend.code(a.dtype), backend.ptr(mask),
                                       backend.ptr(a), backend.ptr(b), backend.ptr(out), C.c_int64(Cn), C.c_int64(d)))
        return out
    raise TypeError("select: float32/float64 only")


def maybe_update_proposal(do_accept, proposal, new_proposal):
    """reference proposals.py:137-174."""
    lib = _lib.load()
    dev = proposal.weight.device
    Cn = proposal.weight.shape[0]
    mask = backend.as_device(do_accept, torch.uint8, dev).expand(Cn).contiguous()
    w = torch.empty(Cn, dtype=torch.float64, device=dev)
    s = torch.empty(Cn, dtype=torch.float64, device=dev)
    # logaddexp of the weights / sum_log_p_accept: the sampling kernel with a dummy uniform
    u = torch.zeros(Cn, dtype=torch.float64, device=dev)
    scratch = torch.empty(Cn, dtype=torch.uint8, device=dev)
    _lib.check(lib.b2h_progressive_sampling(backend.context(dev), 0, backend.ptr(proposal.weight.contiguous()),
                                            backend.ptr(new_proposal.weight.contiguous()),
                                            backend.ptr(proposal.sum_log_p_accept.contiguous()),
                                            backend.ptr(new_proposal.sum_log_p_accept.contiguous()), backend.ptr(u),
                                            backend.ptr(scratch), backend.ptr(w), backend.ptr(s), C.c_int64(Cn)))
    st = IntegratorState(*[_select(mask, n, o) for n, o in zip(new_proposal.state, proposal.state)])
    return ProposalState(st, _select(mask, new_proposal.energy, proposal.energy), w, s)


def _progressive(biased, srng, proposal, new_proposal):
    lib = _lib.load()
    dev = proposal.weight.device
    Cn = proposal.weight.shape[0]
    u = srng.uniform(Cn, dev)
    acc = torch.empty(Cn, dtype=torch.uint8, device=dev)
    w = torch.empty(Cn, dtype=torch.float64, device=dev)
    s = torch.empty(Cn, dtype=torch.float64, device=dev)
    _lib.check(lib.b2h_progressive_sampling(backend.context(dev), int(biased), backend.ptr(proposal.weight.contiguous()),
                                            backend.ptr(new_proposal.weight.contiguous()),
                                            backend.ptr(proposal.sum_log_p_accept.contiguous()),
                                            backend.ptr(new_proposal.sum_log_p_accept.contiguous()), backend.ptr(u),
                                            backend.ptr(acc), backend.ptr(w), backend.ptr(s), C.c_int64(Cn)))
    st = IntegratorState(*[_select(acc, n, o) for n, o in zip(new_proposal.state, proposal.state)])
    return ProposalState(st, _select(acc, new_proposal.energy, proposal.energy), w, s)


def progressive_uniform_sampling(srng, proposal, new_proposal):
    """reference proposals.py:72-102: accept the new proposal with probability expit(w_new - w_old)."""
    return _progressive(0, srng, proposal, new_proposal)


def progressive_biased_sampling(srng, proposal, new_proposal):
    """reference proposals.py:105-134: accept with probability min(1, exp(w_new - w_old))."""
    return _progressive(1, srng, proposal, new_proposal)
