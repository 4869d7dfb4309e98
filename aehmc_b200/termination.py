"""Iterative U-turn criterion (reference termination.py) as batched primitives.

Inside ``nuts.new_kernel`` these steps are fused into the tick engine; the functions here expose the
same three closures for composition and testing.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch

from . import _lib, backend


class TerminationState(NamedTuple):        # reference termination.py:12-16
    momentum_checkpoints: torch.Tensor      # [C, max_num_doublings, d]
    momentum_sum_checkpoints: torch.Tensor
    min_index: torch.Tensor                 # [C] int64
    max_index: torch.Tensor


def iterative_uturn(is_turning_fn):
    """reference termination.py:19-189; ``is_turning_fn`` is the third closure of ``gaussian_metric``."""
    metric = is_turning_fn.metric
    lib = _lib.load()

    def new_state(position, max_num_doublings):
        Cn, d = position.shape
        z = torch.zeros((Cn, int(max_num_doublings), d), dtype=position.dtype, device=position.device)
        i = torch.zeros(Cn, dtype=torch.int64, device=position.device)
        return TerminationState(z, z.clone(), i, i.clone())

    def update(state, momentum_sum, momentum, step):
        mck, sck = state.momentum_checkpoints.clone(), state.momentum_sum_checkpoints.clone()
        imin, imax = state.min_index.clone(), state.max_index.clone()
        Cn, maxd, d = mck.shape
        dev = mck.device
        step = torch.as_tensor(step, dtype=torch.int64, device=dev).expand(Cn).contiguous()
        ms = backend.as_device(momentum_sum, mck.dtype, dev)
        m = backend.as_device(momentum, mck.dtype, dev)
        _lib.check(lib.b2h_termination_update(backend.context(dev), backend.code(mck.dtype), backend.ptr(mck),
                                              backend.ptr(sck), backend.ptr(imin), backend.ptr(imax), backend.ptr(ms),
                                              backend.ptr(m), backend.ptr(step), C.c_int64(Cn), C.c_int64(d),
                                              C.c_int32(maxd)))
        return TerminationState(mck, sck, imin, imax)

    def is_iterative_turning(state, momentum_sum, momentum):
        mck = state.momentum_checkpoints
        Cn, maxd, d = mck.shape
        dev = mck.device
        out = torch.empty(Cn, dtype=torch.uint8, device=dev)
        ms = backend.as_device(momentum_sum, mck.dtype, dev)
        m = backend.as_device(momentum, mck.dtype, dev)
        mt = metric.struct()
        _lib.check(lib.b2h_is_iterative_turning(backend.context(dev), C.byref(mt), backend.code(mck.dtype),
                                                backend.ptr(mck), backend.ptr(state.momentum_sum_checkpoints),
                                                backend.ptr(state.min_index), backend.ptr(state.max_index),
                                                backend.ptr(ms), backend.ptr(m), backend.ptr(out), C.c_int64(Cn),
                                                C.c_int64(d), C.c_int32(maxd)))
        return out.bool()

    return new_state, update, is_iterative_turning


def _find_storage_indices(step):
    """reference termination.py:192-235: (idx_min, idx_max) for each step (popcount / trailing ones)."""
    lib = _lib.load()
    dev = backend.device(step.device if isinstance(step, torch.Tensor) and step.is_cuda else None)
    step = torch.as_tensor(step, dtype=torch.int64, device=dev).reshape(-1).contiguous()
    imin, imax = torch.empty_like(step), torch.empty_like(step)
    _lib.check(lib.b2h_find_storage_indices(backend.context(dev), backend.ptr(step), backend.ptr(imin),
                                            backend.ptr(imax), C.c_int64(step.numel())))
    return imin, imax
