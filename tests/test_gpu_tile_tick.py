"""The tile tick kernel of the per-tick (split) engine (csrc/engine_tile.inl) in each of its layouts: 1 / 8 / 32 chains
per warp (the latter two read their rows through the asynchronous shared-memory ring), 4 / 8 warps per chain, unaligned
rows -- against the oracle under injected draws (bar: tests/test_gpu_nuts.py), and against each other with the native RNG.

B2H_TILE_TC / B2H_TILE_WPC force a layout (the library reads them at every call); B2H_TILE_TICK=0 selects the
register-front tick kernel the tile kernel replaced."""
import numpy as np
import pytest

import parity
from oracle import models as o_models

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ab(cuda_device):
    import aehmc_b200
    return aehmc_b200


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _corr_case(rng, d):
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.1 * np.eye(d)
    return rng.standard_normal(d), cov, np.linalg.inv(cov)


def _run_injected(ab, model, imm, q0, eps, draws, T):
    from aehmc_b200 import _engine
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    info, extras = _engine.run("nuts", model, imm, srng, ab.nuts.new_state(q0, model), torch.as_tensor(eps, dtype=torch.float64),
                               n_transitions=T, store_draws=T)
    return dict(q=_np(info.state.position), p=_np(info.state.momentum), U=_np(info.state.potential_energy),
                g=_np(info.state.potential_energy_grad), acceptance_probability=_np(info.acceptance_probability),
                num_doublings=_np(info.num_doublings), is_turning=_np(info.is_turning),
                is_diverging=_np(info.is_diverging), n_leapfrog=_np(extras["n_leapfrog"]), draws=_np(extras["draws"]))


@pytest.mark.parametrize("tc", [1, 8, 32])
@pytest.mark.parametrize("case", ["corr_dense_d32", "corr_diag_d64", "corr_scalar_d12", "logistic_d8"])
def test_tile_layouts_match_the_oracle(ab, monkeypatch, tc, case):
    """Chains per warp x metric family; 70 chains: the last tile is partial, transitions end on different ticks."""
    monkeypatch.setenv("B2H_TILE_TC", str(tc))
    rng = np.random.default_rng(900 + len(case))
    C, T = 70, 2
    if case.startswith("corr"):
        d = int(case.rsplit("_d", 1)[1])
        mu, cov, prec = _corr_case(rng, d)
        imm = {"dense": cov, "diag": np.diag(cov).copy(), "scalar": np.array(0.7)}[case.split("_")[1]]
        o_model, model = o_models.CorrelatedGaussian(mu, prec), ab.models.CorrelatedGaussian(mu, prec)
        q0 = mu + rng.standard_normal((C, d))
        eps = 0.25 * np.exp(0.2 * rng.standard_normal(C))
    else:
        N, d = 200, 8
        X = np.round(rng.standard_normal((N, d)) * 16) / 16
        y = (rng.random(N) < 0.5).astype(np.float64)
        o_model, model = o_models.LogisticRegression(X, y, 1.0), ab.models.LogisticRegression(X, y, 1.0)
        imm = np.full(d, 4.0 / N)
        q0 = 0.1 * rng.standard_normal((C, d))
        eps = 0.4 * np.exp(0.2 * rng.standard_normal(C))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_model, q0, eps, imm, draws, T)
    got = _run_injected(ab, model, imm, q0, eps, draws, T)
    parity.assert_nuts_parity(got, ref, rtol=1e-9, atol=1e-11, what=f"{case} tc={tc}")
    np.testing.assert_allclose(got["draws"][-1], ref["q"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("wpc", [1, 4, 8])
def test_warps_per_chain_match_the_oracle(ab, monkeypatch, wpc):
    """Long rows (d = 300, dense metric): one chain per warp or per CTA of 4 / 8 warps."""
    monkeypatch.setenv("B2H_TILE_WPC", str(wpc))
    rng = np.random.default_rng(77)
    d, C, T = 300, 5, 1
    mu, cov, prec = _corr_case(rng, d)
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.CorrelatedGaussian(mu, prec), q0, 0.2, cov, draws, T)
    got = _run_injected(ab, ab.models.CorrelatedGaussian(mu, prec), cov, q0, 0.2, draws, T)
    parity.assert_nuts_parity(got, ref, rtol=1e-9, atol=1e-11, what=f"wpc={wpc}")


@pytest.mark.parametrize("tc", [8, 32])
def test_window_adaptation_through_the_ring(ab, monkeypatch, tc):
    """Per-chain diagonal metric adapted in the tick engine: its rows travel through the ring like the state's."""
    monkeypatch.setenv("B2H_TILE_TC", str(tc))
    rng = np.random.default_rng(23)
    C, W, N, d = 40, 22, 96, 4
    X = rng.standard_normal((N, d))
    y = (rng.random(N) < 0.5).astype(np.float64)
    q0 = 0.5 * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W, d)
    ref = parity.oracle_nuts(o_models.LogisticRegression(X, y, 1.0), q0, 1.0, np.ones(d), draws, W, schedule_steps=W)
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    model = ab.models.LogisticRegression(X, y, 1.0)
    kernel = ab.nuts.new_kernel(srng, model)
    state, (step_size, imm_out), _ = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W)
    np.testing.assert_allclose(_np(step_size), ref["eps"], rtol=1e-6)
    np.testing.assert_allclose(_np(imm_out), ref["imm"], rtol=1e-6)
    np.testing.assert_allclose(_np(state.position), ref["q"], rtol=1e-5, atol=1e-8)


def _philox_run(ab, model, imm, q0, eps, T, seed=5):
    from aehmc_b200 import _engine
    info, extras = _engine.run("nuts", model, imm, ab.RandomStream(seed=seed), ab.nuts.new_state(q0, model), eps,
                               n_transitions=T)
    return _np(info.state.position), _np(info.num_doublings), _np(extras["n_leapfrog"])


@pytest.mark.parametrize("dtype_name", ["float32", "float64"])
def test_layouts_agree_bit_for_bit_with_the_native_rng(ab, monkeypatch, dtype_name):
    """3001 chains x 6 transitions with Philox draws: 1, 8 and 32 chains per warp add the same partial sums in the same
    order, so positions are IDENTICAL; the register-front kernel (another summation order) agrees in tree shapes for
    nearly all chains and in positions to rounding."""
    dtype = getattr(torch, dtype_name)
    rng = np.random.default_rng(31)
    N, d, C, T = 256, 64 if dtype_name == "float64" else 128, 3001, 6
    X = np.round(rng.standard_normal((N, d)) * 8) / 8
    y = (rng.random(N) < 0.5).astype(np.float64)
    model = ab.models.LogisticRegression(X, y, 1.0, dtype=dtype)
    imm = np.full(d, 4.0 / N)
    q0 = 0.1 * rng.standard_normal((C, d))
    res = {}
    for tc in (1, 8, 32):
        monkeypatch.setenv("B2H_TILE_TC", str(tc))
        res[tc] = _philox_run(ab, model, imm, q0, 0.35, T)
    for tc in (8, 32):
        np.testing.assert_array_equal(res[tc][1], res[1][1])
        np.testing.assert_array_equal(res[tc][2], res[1][2])
        np.testing.assert_array_equal(res[tc][0], res[1][0])
    monkeypatch.delenv("B2H_TILE_TC")
    monkeypatch.setenv("B2H_TILE_TICK", "0")
    q_old, nd_old, nl_old = _philox_run(ab, model, imm, q0, 0.35, T)
    same = (nd_old == res[1][1]) & (nl_old == res[1][2])
    assert same.mean() >= (0.97 if dtype_name == "float64" else 0.5)
    if dtype_name == "float64":            # float32 trajectories decorrelate within a few transitions
        np.testing.assert_allclose(q_old[same], res[1][0][same], rtol=1e-6, atol=1e-8)


def test_tick_timer_brackets_the_tick_kernel_and_the_gradient_call(ab):
    """b2h_tick_timer (include/b200hmc.h): mode 1 times every launch of the tick kernel, mode 2 every gradient call, of the
    per-tick engine; a max_ticks run of T ticks has T gradient calls and T - 1 full (post + pre) tick kernels -- the
    post-only launch that closes the call is not counted."""
    import ctypes as C
    from aehmc_b200 import _engine, _lib, backend
    lib = _lib.load()
    rng = np.random.default_rng(5)
    N, d, Cn, ticks = 128, 8, 64, 6
    X = rng.standard_normal((N, d))
    y = (rng.random(N) < 0.5).astype(np.float64)
    model = ab.models.LogisticRegression(X, y, 1.0)
    state = ab.nuts.new_state(0.1 * rng.standard_normal((Cn, d)), model)
    ctx = backend.context(model.device)
    for mode in (1, 2):
        _lib.check(lib.b2h_tick_timer(ctx, mode))
        _engine.run("nuts", model, np.full(d, 4.0 / N), ab.RandomStream(seed=1), state, 0.3, max_ticks=ticks)
        ms, n = C.c_double(0.0), C.c_int64(0)
        _lib.check(lib.b2h_tick_timer_read(ctx, C.byref(ms), C.byref(n)))
        _lib.check(lib.b2h_tick_timer(ctx, 0))
        assert n.value == (ticks - 1 if mode == 1 else ticks) and ms.value > 0.0
    # off: nothing is recorded
    _engine.run("nuts", model, np.full(d, 4.0 / N), ab.RandomStream(seed=1), state, 0.3, max_ticks=ticks)
    ms, n = C.c_double(0.0), C.c_int64(0)
    _lib.check(lib.b2h_tick_timer_read(ctx, C.byref(ms), C.byref(n)))
    assert n.value == 0
