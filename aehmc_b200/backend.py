"""Device plumbing: one context per (device, stream), raw-pointer helpers.

PyTorch is used for device memory, streams and torch.distributed only; every
computation goes through the C-ABI of libb200hmc.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_contexts = {}


def torch_dtype(dtype):
    if dtype in (torch.float64, "float64", "f64", np.float64, _lib.F64):
        return torch.float64
    if dtype in (torch.float32, "float32", "f32", np.float32, _lib.F32):
        return torch.float32
    raise ValueError(f"dtype must be float32 or float64, got {dtype!r}")


def code(dtype):
    return _lib.F64 if torch_dtype(dtype) == torch.float64 else _lib.F32


def device(dev=None):
    if not torch.cuda.is_available():
        raise _lib.B200HMCError("no CUDA device: aehmc_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(dev)
    if dev.type != "cuda":
        raise _lib.B200HMCError("aehmc_b200 tensors must live on a CUDA device")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def context(dev):
    """b2h_ctx bound to torch's CURRENT stream of ``dev`` (so CUDA events recorded through torch
    bracket the library's kernels)."""
    dev = device(dev)
    stream = torch.cuda.current_stream(dev)
    key = (dev.index, stream.cuda_stream)
    ctx = _contexts.get(key)
    if ctx is None:
        lib = _lib.load()
        handle = C.c_void_p()
        _lib.check(lib.b2h_ctx_create(C.c_int(dev.index), C.c_void_p(stream.cuda_stream), C.byref(handle)))
        ctx = handle
        _contexts[key] = ctx
    return ctx


def as_device(x, dtype, dev):
    """Upload (numpy / python / tensor) to a contiguous device tensor of ``dtype``."""
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=dev).contiguous()


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class Workspace:
    """Grow-only device scratch owned by the caller side of the C-ABI."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, dev):
        nbytes = int(nbytes)
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != dev:
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        return self.buf
