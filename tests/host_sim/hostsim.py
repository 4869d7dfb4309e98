"""TEST TOOLING ONLY: ctypes driver of tests/host_sim/sim.cpp (engine.cuh compiled with g++)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "build", "libhostsim.so")
SRC = os.path.join(ROOT, "tests", "host_sim", "sim.cpp")


def build():
    deps = [SRC, os.path.join(ROOT, "tests/host_sim/cuda_shim.h")] + [
        os.path.join(ROOT, "aehmc_b200/csrc", f) for f in ("engine.cuh", "models.cuh", "common.cuh")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", SO, SRC])
    return SO


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def run(model_kind, a, b, s0, imm, q, eps, draws, n_transitions, *, hmc_L=0, maxd=10, div_thr=1000.0,
        n_store=0, schedule=None, target=0.8, init_step_size=1.0, reg_front=False, exact_doubling=False):
    """draws: dict of arrays [C, T, ...].  Returns dict of outputs.  reg_front: keep the integration front
    in "registers" (engine.cuh RegFront) instead of the edge arrays."""
    lib = ctypes.CDLL(build())
    lib.sim_set_reg_front(ctypes.c_int(16 if reg_front else 0))
    lib.sim_set_exact_doubling(ctypes.c_int(1 if exact_doubling else 0))
    q = np.ascontiguousarray(q, dtype=np.float64).copy()
    C, d = q.shape
    imm = np.asarray(imm, dtype=np.float64)
    if imm.ndim == 0:
        imm_kind, imm_arr, imm_scalar = 0, None, float(imm)
    elif imm.ndim == 1:
        imm_kind, imm_arr, imm_scalar = 1, np.ascontiguousarray(imm), 0.0
    else:
        imm_kind, imm_arr, imm_scalar = 2, np.ascontiguousarray(imm), 0.0
    eps = np.ascontiguousarray(np.broadcast_to(np.asarray(eps, dtype=np.float64), (C,))).copy()
    p = np.zeros_like(q); g = np.zeros_like(q); U = np.zeros(C)
    acc = np.zeros(C); nd = np.zeros(C, np.int32); turning = np.zeros(C, np.uint8); div = np.zeros(C, np.uint8)
    nleap = np.zeros(C, np.int32)
    store = np.zeros((max(n_store, 1), C, d))
    z = np.ascontiguousarray(draws["z"], dtype=np.float64)
    n_inj = z.shape[1]
    g_ = lambda k: None if draws.get(k) is None else np.ascontiguousarray(draws[k], dtype=np.float64)
    u_dir, u_b, u_u, u_a = g_("u_dir"), g_("u_biased"), g_("u_uniform"), g_("u_accept")
    a = None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    adapt_steps, stage, wend = 0, None, None
    if schedule is not None:
        adapt_steps = len(schedule)
        stage = np.array([s for s, _ in schedule], dtype=np.uint8)
        wend = np.array([e for _, e in schedule], dtype=np.uint8)
    imm_out = np.zeros_like(q)
    u8 = ctypes.c_ubyte
    lib.sim_run.restype = ctypes.c_int
    rc = lib.sim_run(
        ctypes.c_int(model_kind), ctypes.c_int(1 if hmc_L > 0 else 0), ctypes.c_int(C), ctypes.c_int(d),
        ctypes.c_int(maxd), _p(a), _p(b), ctypes.c_double(s0), ctypes.c_int(imm_kind), _p(imm_arr),
        ctypes.c_double(imm_scalar), _p(q), _p(p), _p(U), _p(g), _p(eps), ctypes.c_longlong(n_inj), _p(z),
        _p(u_dir), _p(u_b), _p(u_u), _p(u_a), ctypes.c_int(n_transitions), ctypes.c_int(hmc_L),
        ctypes.c_double(div_thr), _p(acc), _p(nd, ctypes.c_int), _p(turning, u8), _p(div, u8),
        _p(nleap, ctypes.c_int), _p(store), ctypes.c_int(n_store), ctypes.c_int(adapt_steps), _p(stage, u8),
        _p(wend, u8), ctypes.c_double(target), ctypes.c_double(init_step_size), _p(imm_out))
    assert rc == 0
    return dict(q=q, p=p, U=U, g=g, eps=eps, acceptance_probability=acc, num_doublings=nd, is_turning=turning,
                is_diverging=div, n_leapfrog=nleap, draws=store[:n_store], imm=imm_out)
