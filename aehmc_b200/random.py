"""The ``srng`` argument of the kernels.

``RandomStream(seed)`` is the native mode: Philox4x32-10 keyed by (seed, global chain id,
transition, draw kind, slot), so results do not depend on how chains are sharded over GPUs.
``InjectedDraws`` is the validation mode of BASELINE.json's north star: index-addressed
normals / uniforms (the reference's own draws, or any others) decide every random choice,
which makes tree depth, step counts, divergence flags and accepted states comparable
bit-for-bit with the reference semantics (oracle/streams.py).
"""
from __future__ import annotations

import torch

from . import _lib, backend


class RandomStream:
    def __init__(self, seed=0, chain_offset=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.chain_offset = int(chain_offset)
        self.transition = 0           # advanced by the kernels (the reference returns `updates`)

    def struct(self, n_transitions=1):
        return _lib.Rng(_lib.RNG_PHILOX, 0, self.seed, self.chain_offset, self.transition, 0,
                        None, None, None, None, None), ()

    def advance(self, n):
        self.transition += int(n)

    def uniform(self, n, device):
        """[n] float64 uniforms for the stand-alone proposal samplers (the accept slot of one Philox transition)."""
        import ctypes as C
        lib = _lib.load()
        dev = backend.device(device)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.check(lib.b2h_philox_fill(backend.context(dev), C.c_uint64(self.seed), C.c_uint64(self.chain_offset),
                                       C.c_uint64(self.transition), C.c_int64(n), C.c_int64(1), C.c_int64(1),
                                       C.c_int32(1), None, None, None, None, backend.ptr(out)))
        self.advance(1)
        return out


class InjectedDraws:
    """z [C,T,d], u_dir [C,T,max], u_biased [C,T,max], u_uniform [C,T,2**max-1], u_accept [C,T] (float64)."""

    def __init__(self, z, u_dir=None, u_biased=None, u_uniform=None, u_accept=None, device=None):
        dev = backend.device(device)
        up = lambda a: None if a is None else backend.as_device(a, torch.float64, dev)
        self.z, self.u_dir, self.u_biased = up(z), up(u_dir), up(u_biased)
        self.u_uniform, self.u_accept = up(u_uniform), up(u_accept)
        first = next(a for a in (self.z, self.u_dir, self.u_biased, self.u_uniform, self.u_accept) if a is not None)
        self.n_injected = int(first.shape[1])
        self.transition = 0
        self._cursor = 0

    def struct(self, n_transitions=1):
        if self.transition + int(n_transitions) > self.n_injected:
            raise ValueError(f"injected draws exhausted: {self.n_injected} transitions injected, "
                             f"{self.transition} consumed, {n_transitions} requested")
        keep = (self.z, self.u_dir, self.u_biased, self.u_uniform, self.u_accept)
        p = lambda t: None if t is None else t.data_ptr()
        return _lib.Rng(_lib.RNG_INJECTED, 0, 0, 0, self.transition, self.n_injected, p(self.z), p(self.u_dir),
                        p(self.u_biased), p(self.u_uniform), p(self.u_accept)), keep

    def advance(self, n):
        self.transition += int(n)

    def uniform(self, n, device):
        """stand-alone proposal samplers: the k-th call consumes u_accept[:, k]."""
        if self.u_accept is None or self._cursor >= self.u_accept.shape[1]:
            raise ValueError("injected draws: u_accept exhausted")
        out = self.u_accept[:, self._cursor].contiguous()
        self._cursor += 1
        return out
