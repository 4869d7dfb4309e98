"""Pooled (cross-chain) window adaptation on the GPU: the Welford kernels against the reference's recurrence, and the
whole pooled warm-up (diagonal and dense inverse mass matrix, many chains) against oracle.adaptation.run_pooled."""
import ctypes as C

import numpy as np
import pytest

import parity
from oracle import adaptation as o_adapt
from oracle import kernels as o_kernels
from oracle import models as o_models

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ab(cuda_device):
    import aehmc_b200
    return aehmc_b200


def _welford(values, full):
    init, update, _ = o_adapt.welford_covariance(full)
    mean, m2, n = init(values.shape[1])
    for v in values:
        mean, m2, n = update(v, mean, m2, n)
    return n, mean, m2


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("full", [False, True])
@pytest.mark.parametrize("T, Cn, d", [(3, 7, 5), (4, 33, 10), (2, 129, 130)])
def test_pooled_update_matches_the_welford_recurrence(ab, T, Cn, d, full, dt):
    from aehmc_b200.mass_matrix import PooledWelford
    rng = np.random.default_rng(T + Cn + d)
    scale, shift = np.exp(rng.standard_normal(d)), 10.0 * rng.standard_normal(d)
    x = torch.tensor(rng.standard_normal((2 * T, Cn, d)) * scale + shift, dtype=getattr(torch, dt), device="cuda")
    pool = PooledWelford(d, full, "cuda")
    pool.update(x[:T])
    pool.update(x[T:])                                   # second block: Chan's merge with the running state
    n, mean, m2 = _welford(x.double().cpu().numpy().reshape(-1, d), full)
    assert pool.n == n
    np.testing.assert_allclose(pool.mean.cpu().numpy(), mean, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(pool.m2.cpu().numpy(), m2, rtol=1e-10, atol=1e-9 * np.abs(m2).max())
    _, _, final = o_adapt.covariance_adaptation(full)
    np.testing.assert_allclose(pool.final().cpu().numpy(), final((mean, m2, n)), rtol=1e-10, atol=1e-12)


def test_welford_merge_entry_point(ab):
    from aehmc_b200 import _lib, backend
    rng = np.random.default_rng(3)
    x = rng.standard_normal((50, 6)) + 3.0
    for full in (False, True):
        a, b = _welford(x[:20], full), _welford(x[20:], full)
        up = lambda v: torch.tensor(np.asarray(v), dtype=torch.float64, device="cuda")
        mean_a, m2_a, mean_b, m2_b = up(a[1]), up(a[2]), up(b[1]), up(b[2])
        scratch = torch.empty(6, dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().b2h_welford_merge(backend.context(mean_a.device), C.c_int64(6), C.c_int32(int(full)),
                                                 C.c_int64(a[0]), backend.ptr(mean_a), backend.ptr(m2_a), C.c_int64(b[0]),
                                                 backend.ptr(mean_b), backend.ptr(m2_b), backend.ptr(scratch)))
        n, mean, m2 = _welford(x, full)
        np.testing.assert_allclose(mean_a.cpu().numpy(), mean, rtol=1e-13)
        np.testing.assert_allclose(m2_a.cpu().numpy(), m2, rtol=1e-11, atol=1e-10)


def _oracle_pooled(model, q0, draws, W, full):
    Cn = q0.shape[0]
    ks = [o_kernels.nuts_new_kernel(parity.chain_draws(draws, c), model) for c in range(Cn)]
    sts = [o_kernels.new_state(q0[c].copy(), model) for c in range(Cn)]
    with np.errstate(all="ignore"):
        states, (eps, imm), _ = o_adapt.run_pooled(ks, sts, W, is_mass_matrix_full=full)
    return np.stack([s.position for s in states]), eps, imm


@pytest.mark.parametrize("chunk_bytes", [2 << 30, 400])            # 400 bytes: at most one transition per engine call
def test_pooled_warmup_diagonal_matches_oracle(ab, chunk_bytes):
    """fused engine, shared diagonal metric re-estimated from ALL chains at the window ends."""
    rng = np.random.default_rng(23)
    Cn, W, d = 7, 24, 5
    mu, sigma = rng.standard_normal(d), np.exp(rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((Cn, d))
    draws = parity.random_draws(rng, Cn, W, d)
    q_ref, eps_ref, imm_ref = _oracle_pooled(o_models.IIDGaussian(mu, sigma), q0, draws, W, False)
    model = ab.models.IIDGaussian(mu, sigma)
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    kernel = ab.nuts.new_kernel(srng, model)
    state, (eps, imm), info = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W, pooled=True,
                                                       max_chunk_bytes=chunk_bytes)
    assert imm.shape == (d,) and eps.shape == (Cn,) and len(info["pooled_window_sizes"]) >= 1
    np.testing.assert_allclose(imm.cpu().numpy(), imm_ref, rtol=1e-7)
    np.testing.assert_allclose(eps.cpu().numpy(), eps_ref, rtol=1e-7)
    np.testing.assert_allclose(state.position.cpu().numpy(), q_ref, rtol=1e-6, atol=1e-9)


def test_pooled_warmup_dense_matches_oracle(ab):
    """is_mass_matrix_full=True with MANY chains (impossible per chain: a dense metric is shared): split engine,
    correlated Gaussian, the dense metric rebuilt from the pooled covariance at the window ends."""
    rng = np.random.default_rng(29)
    Cn, W, d = 9, 25, 4
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.3 * np.eye(d)
    mu, prec = rng.standard_normal(d), np.linalg.inv(cov)
    q0 = mu + rng.standard_normal((Cn, d))
    draws = parity.random_draws(rng, Cn, W, d)
    q_ref, eps_ref, imm_ref = _oracle_pooled(o_models.CorrelatedGaussian(mu, prec), q0, draws, W, True)
    model = ab.models.CorrelatedGaussian(mu, prec)
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    kernel = ab.nuts.new_kernel(srng, model)
    state, (eps, imm), _ = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W, pooled=True,
                                                    is_mass_matrix_full=True)
    assert imm.shape == (d, d)
    np.testing.assert_allclose(imm.cpu().numpy(), imm_ref, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(eps.cpu().numpy(), eps_ref, rtol=1e-6)
    np.testing.assert_allclose(state.position.cpu().numpy(), q_ref, rtol=1e-5, atol=1e-8)


def test_pooled_warmup_adapts_a_dense_metric_for_many_chains(ab):
    """Statistical check in native-RNG mode: 2048 chains on a correlated Gaussian recover the covariance."""
    rng = np.random.default_rng(31)
    Cn, W, d = 2048, 150, 16
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.1 * np.eye(d)
    model = ab.models.CorrelatedGaussian(np.zeros(d), np.linalg.inv(cov))
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=4), model)
    q0 = rng.standard_normal((Cn, d))
    state, (eps, imm), info = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W, pooled=True,
                                                       is_mass_matrix_full=True)
    rel = np.abs(imm.cpu().numpy() - cov).max() / np.abs(cov).max()
    assert rel < 0.05, rel
    assert 0.2 < float(eps.median()) < 3.0
