"""covariance_adaptation (reference mass_matrix.py:12-120) over batched chains."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, algorithms, backend


def covariance_adaptation(is_mass_matrix_full=False):
    wc_init, wc_update, wc_final = algorithms.welford_covariance(is_mass_matrix_full)
    lib = _lib.load()
    full = 1 if is_mass_matrix_full else 0

    def init(n_dims, num_chains=1, dtype=torch.float64, device=None):
        dev = backend.device(device)
        if is_mass_matrix_full:
            imm = torch.eye(n_dims, dtype=dtype, device=dev).expand(num_chains, n_dims, n_dims).contiguous()
        else:
            imm = torch.ones((num_chains, n_dims), dtype=dtype, device=dev)
        return imm, wc_init(n_dims, num_chains, dtype, dev)

    def update(position, wc_state):
        return wc_update(position, *wc_state)

    def final(wc_state):
        """Stan's shrinkage (mass_matrix.py:103-116): (n/(n+5)) cov + 1e-3 (5/(n+5)) [* I]."""
        _, m2, n = wc_state
        out = torch.empty_like(m2)
        Cn, d = m2.shape[0], m2.shape[1]
        _lib.check(lib.b2h_mass_matrix_final(backend.context(m2.device), backend.code(m2.dtype), backend.ptr(m2),
                                             backend.ptr(n), backend.ptr(out), C.c_int64(Cn), C.c_int64(d),
                                             C.c_int32(full)))
        return out

    return init, update, final
