"""Micro-benchmark of b2h_dense_apply against cuBLAS on the c2 shape (development tool)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from aehmc_b200 import _lib, backend
lib = _lib.load()
dev = torch.device("cuda:0")
for (Cn, d) in ((4096, 1000), (4608, 1000), (4096, 1024), (8192, 1000)):
    a = torch.randn((Cn, d), dtype=torch.float64, device=dev); m = torch.randn((d, d), dtype=torch.float64, device=dev)
    out = torch.empty_like(a); ctx = backend.context(dev)
    def run():
        _lib.check(lib.b2h_dense_apply(ctx, _lib.F64, backend.ptr(a), backend.ptr(m), backend.ptr(out), C.c_int64(Cn), C.c_int64(d)))
    for f, name in ((run, "b2h"), (lambda: torch.matmul(a, m), "cublas")):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name:7s} C={Cn} d={d}: {ms:.4f} ms  {2.0*Cn*d*d/ms/1e9:.2f} TFLOP/s  env={os.environ.get('B2H_GEMM_ASYNC','')},{os.environ.get('B2H_GEMM_BN','')},{os.environ.get('B2H_DMMA_K','')}")
    err = (out - a @ m).abs().max().item()
    print("   max abs err vs cublas", err)
