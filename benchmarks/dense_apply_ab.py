"""b2h_dense_apply at the c2 shape (4096 x 1000 x 1000, float64): TFLOP/s of the hand-written DMMA kernel (variant selected
by B2H_GEMM_WARPS / B2H_GEMM_BN) next to cuBLAS on the same shape, plus a correctness check against torch.matmul."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aehmc_b200 import _lib, backend  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
for Cn, d in ((4096, 1000), (4096, 1024), (8192, 1000)):
    a = torch.randn((Cn, d), dtype=torch.float64, device=dev)
    m = torch.randn((d, d), dtype=torch.float64, device=dev)
    m = m + m.T
    out = torch.empty_like(a)
    ctx = backend.context(dev)
    fn = lambda: _lib.check(lib.b2h_dense_apply(ctx, _lib.F64, backend.ptr(a), backend.ptr(m), backend.ptr(out), C.c_int64(Cn), C.c_int64(d)))

    def ms(f, reps):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t = ms(fn, 50)
    ref = a @ m
    err = float((out - ref).abs().max() / ref.abs().max())
    tc = ms(lambda: torch.matmul(a, m), 50)
    fl = 2.0 * Cn * d * d
    print(f"warps={os.environ.get('B2H_GEMM_WARPS', 'default')} bn={os.environ.get('B2H_GEMM_BN', 'auto')} {Cn}x{d}x{d}: "
          f"{t * 1e3:.1f} us = {fl / t / 1e9:.2f} TFLOP/s; cuBLAS {tc * 1e3:.1f} us = {fl / tc / 1e9:.2f} TFLOP/s; rel err {err:.1e}", flush=True)
