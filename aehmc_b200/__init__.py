"""aehmc_b200 -- B200-native many-chain HMC/NUTS behind the aehmc function surface.

Module and function names follow aesara-devs/aehmc (hmc, nuts, integrators, metrics, termination,
trajectory, algorithms, step_size, mass_matrix, window_adaptation).  Positions are batched
``[chains, dim]`` CUDA tensors, ``logprob_fn`` is a model descriptor from ``aehmc_b200.models`` and
``srng`` is ``RandomStream(seed)`` (Philox) or ``InjectedDraws`` (validation mode).  All computation
happens in libb200hmc.so (hand-written sm_100a CUDA) through its C-ABI; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .random import InjectedDraws, RandomStream  # noqa: F401
from . import (algorithms, diagnostics, hmc, integrators, mass_matrix, metrics, models, nuts, sampling, step_size,  # noqa: F401
               termination, trajectory, utils, window_adaptation, proposals)

__version__ = "0.1.0"
