"""GPU parity of the HMC/NUTS transition kernels (validation mode: injected draws).

Bar (BASELINE.json north_star): tree depth, number of integration steps, divergence / turning flags
bit-exact; positions, momenta, gradients, energies within 1e-10 relative in float64 and 1e-4 in
float32, per transition.  The oracle is pinned to the reference by tests/test_oracle_golden.py."""
import numpy as np
import pytest

import parity
from oracle import adaptation as o_adapt
from oracle import kernels as o_kernels
from oracle import models as o_models
from oracle import streams as o_streams

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ab(cuda_device):
    import aehmc_b200
    return aehmc_b200


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def gpu_nuts(ab, model, imm, q0, eps, draws, T, maxd=10, group=0, schedule=None, div_thr=1000.0):
    from aehmc_b200 import _engine
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    state = ab.nuts.new_state(q0, model)
    Cn = q0.shape[0]
    adapt = None
    if schedule is not None:
        adapt = _engine.AdaptState(Cn, schedule, model.device)
    info, extras = _engine.run("nuts", model, imm, srng, state, torch.as_tensor(eps, dtype=torch.float64), n_transitions=T,
                               max_num_expansions=maxd, divergence_threshold=div_thr, store_draws=T, group=group,
                               adapt=adapt)
    out = dict(q=_np(info.state.position), p=_np(info.state.momentum), U=_np(info.state.potential_energy),
               g=_np(info.state.potential_energy_grad), acceptance_probability=_np(info.acceptance_probability),
               num_doublings=_np(info.num_doublings), is_turning=_np(info.is_turning),
               is_diverging=_np(info.is_diverging), n_leapfrog=_np(extras["n_leapfrog"]), draws=_np(extras["draws"]),
               stats=_np(extras["draw_stats"]), eps=_np(extras["step_size"]),
               imm=_np(extras["inverse_mass_matrix"]))
    return out


def test_readme_quickstart_bit_exact_on_gpu(ab):
    """reference README.md:22-55: seed 0, N(0,1), q0 = 0, step 1e-2, imm = 1.0 -> 1.1034719409361107.
    The reference's own RNG streams are replayed as injected draws; the GPU position is bit-identical."""
    rec = o_streams.RecordingStreamDraws(0, "nuts")
    om = o_models.IIDGaussian([0.0], [1.0], const=o_models._LOG_SQRT_2PI)
    info, extras = o_kernels.nuts_new_kernel(rec, om)(o_kernels.new_state(np.zeros(1), om), 1e-2, np.float64(1.0))
    assert float(info.state.position[0]) == 1.1034719409361107
    inj = {k: v[None] for k, v in rec.injected().items()}
    model = ab.models.IIDGaussian([0.0], [1.0], const=o_models._LOG_SQRT_2PI)
    for group in (1, 8, 32):
        kernel = ab.nuts.new_kernel(ab.InjectedDraws(inj["z"], inj["u_dir"], inj["u_biased"], inj["u_uniform"]), model)
        state = ab.nuts.new_state(np.zeros((1, 1)), model)
        from aehmc_b200 import _engine
        out, ex = _engine.run("nuts", model, 1.0, kernel.spec["srng"], state, 1e-2, group=group)
        assert out.state.position.item() == 1.1034719409361107, group
        assert out.num_doublings.item() == 8 and ex["n_leapfrog"].item() == 136
        assert not out.is_turning.item() and not out.is_diverging.item()
        assert out.acceptance_probability.item() == pytest.approx(info.acceptance_probability, rel=1e-13)
    # and through the public kernel signature
    kernel = ab.nuts.new_kernel(ab.InjectedDraws(inj["z"], inj["u_dir"], inj["u_biased"], inj["u_uniform"]), model)
    chain_info, updates = kernel(ab.nuts.new_state(np.zeros((1, 1)), model), 1e-2, 1.0)
    assert chain_info.state.position.item() == 1.1034719409361107


@pytest.mark.parametrize("d, imm_kind, eps, group", [
    (1, "scalar", 0.3, 0), (5, "diag", 0.4, 1), (5, "diag", 1.7, 8), (3, "per_chain", 0.2, 0),
    (7, "scalar", 1e-3, 0), (2, "diag", 30.0, 0), (40, "diag", 0.3, 0), (40, "per_chain", 0.3, 32),
    (100, "diag", 0.25, 0), (130, "diag", 0.2, 256), (700, "scalar", 0.1, 0),
])
def test_nuts_iid_gaussian(ab, d, imm_kind, eps, group):
    rng = np.random.default_rng(100 + d)
    C, T = (12, 3) if d <= 100 else (5, 2)
    mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
    imm = {"scalar": np.float64(0.7), "diag": sigma ** 2 * np.exp(0.2 * rng.standard_normal(d)),
           "per_chain": np.exp(0.3 * rng.standard_normal((C, d)))}[imm_kind]
    q0 = mu + sigma * rng.standard_normal((C, d))
    eps_c = eps * np.exp(0.3 * rng.standard_normal(C))
    maxd = 6 if eps < 0.01 else 10
    draws = parity.random_draws(rng, C, T, d, maxd)
    om = o_models.IIDGaussian(mu, sigma, const=0.25)
    ref = parity.oracle_nuts(om, q0, eps_c, imm, draws, T, maxd=maxd, per_chain_imm=imm_kind == "per_chain")
    model = ab.models.IIDGaussian(mu, sigma, const=0.25)
    gi = ab.metrics.per_chain(imm) if imm_kind == "per_chain" else (float(imm) if imm_kind == "scalar" else imm)
    got = gpu_nuts(ab, model, gi, q0, eps_c, draws, T, maxd=maxd, group=group)
    parity.assert_nuts_parity(got, ref, rtol=1e-10, what=f"iid d={d} G={group}")
    np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-10, atol=1e-12)
    for c in range(C):
        assert [int(x) for x in got["stats"][:, c, 1]] == [h[0] for h in ref["hist"][c]]
        assert [int(x) for x in got["stats"][:, c, 2]] == [h[1] for h in ref["hist"][c]]


@pytest.mark.parametrize("group", [1, 8])
@pytest.mark.parametrize("eps", [0.05, 0.5, 3.0])
def test_nuts_funnel(ab, eps, group):
    rng = np.random.default_rng(7)
    C, T, d = 40, 3, 10
    q0 = rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.NealFunnel(d), q0, eps, np.ones(d), draws, T)
    got = gpu_nuts(ab, ab.models.NealFunnel(d), np.ones(d), q0, eps, draws, T, group=group)
    parity.assert_nuts_parity(got, ref, rtol=1e-9, what="funnel")
    if eps >= 3.0:
        assert ref["is_diverging"].any()


@pytest.mark.parametrize("group", [1, 8])
def test_nuts_eight_schools(ab, group):
    rng = np.random.default_rng(8)
    C, T, d = 40, 4, 10
    q0 = 0.5 * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.EightSchools(), q0, 0.3, np.ones(d), draws, T)
    got = gpu_nuts(ab, ab.models.EightSchools(), np.ones(d), q0, 0.3, draws, T, group=group)
    parity.assert_nuts_parity(got, ref, rtol=1e-9, what="eight schools")
    assert len(set(ref["num_doublings"].tolist())) > 1


def _corr_case(rng, d):
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.1 * np.eye(d)
    return rng.standard_normal(d), cov, np.linalg.inv(cov)


@pytest.mark.parametrize("d, metric_kind, group", [(6, "dense", 8), (6, "diag", 8), (40, "dense", 32),
                                                   (150, "dense", 0), (600, "dense", 256)])
def test_nuts_correlated_gaussian_split_engine(ab, d, metric_kind, group):
    """Config 2 shape: correlated Gaussian target, dense inverse mass matrix (split tick engine: dense
    applies of all chains as one contraction per half-step)."""
    rng = np.random.default_rng(200 + d)
    C, T = (10, 2) if d <= 150 else (4, 1)
    mu, cov, prec = _corr_case(rng, d)
    imm = cov if metric_kind == "dense" else np.diag(cov).copy()
    q0 = mu + rng.standard_normal((C, d))
    eps = 0.25 * np.exp(0.2 * rng.standard_normal(C))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.CorrelatedGaussian(mu, prec), q0, eps, imm, draws, T)
    got = gpu_nuts(ab, ab.models.CorrelatedGaussian(mu, prec), imm, q0, eps, draws, T, group=group)
    parity.assert_nuts_parity(got, ref, rtol=1e-9, atol=1e-11, what=f"corr d={d} {metric_kind}")


def test_nuts_logistic_regression_split_engine(ab):
    """Config 3 shape (small): Bayesian logistic regression, FP64 FMA gradient path."""
    rng = np.random.default_rng(11)
    N, d, C, T = 300, 8, 12, 2
    X = np.round(rng.standard_normal((N, d)) * 16) / 16
    beta = rng.standard_normal(d) / np.sqrt(d)
    y = (rng.random(N) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    q0 = 0.1 * rng.standard_normal((C, d))
    imm = np.full(d, 4.0 / N)
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.LogisticRegression(X, y, 1.0), q0, 0.4, imm, draws, T)
    got = gpu_nuts(ab, ab.models.LogisticRegression(X, y, 1.0), imm, q0, 0.4, draws, T)
    parity.assert_nuts_parity(got, ref, rtol=1e-9, atol=1e-11, what="logistic")


def test_nuts_float32_within_1e4(ab):
    rng = np.random.default_rng(12)
    C, T, d = 24, 1, 20
    mu, sigma = rng.standard_normal(d), np.exp(0.3 * rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.IIDGaussian(mu, sigma), q0, 0.3, sigma ** 2, draws, T)
    got = gpu_nuts(ab, ab.models.IIDGaussian(mu, sigma, dtype=torch.float32), sigma ** 2, q0, 0.3, draws, T)
    same = (got["num_doublings"] == ref["num_doublings"]) & (got["n_leapfrog"] == ref["n_leapfrog"])
    assert same.mean() >= 0.9          # fp32 rounding may flip a near-tie U-turn test
    for k in ("q", "p", "g"):
        np.testing.assert_allclose(got[k][same], ref[k][same], rtol=1e-4, atol=1e-4)


def test_window_adaptation_on_device(ab):
    """window_adaptation.run (reference window_adaptation.py:17-116) fused into the tick engine."""
    rng = np.random.default_rng(19)
    C, W, d = 6, 22, 5
    mu, sigma = rng.standard_normal(d), np.exp(rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W, d)
    sched = o_adapt.build_schedule(W)
    ref = parity.oracle_nuts(o_models.IIDGaussian(mu, sigma), q0, 1.0, np.ones(d), draws, W, schedule_steps=W)
    model = ab.models.IIDGaussian(mu, sigma)
    imm = ab.metrics.per_chain(torch.ones((C, d), dtype=torch.float64, device="cuda"))
    got = gpu_nuts(ab, model, imm, q0, 1.0, draws, W, schedule=sched)
    parity.assert_nuts_parity(got, ref, rtol=1e-6, what="adapt")
    np.testing.assert_allclose(got["eps"], ref["eps"], rtol=1e-7)
    np.testing.assert_allclose(got["imm"], ref["imm"], rtol=1e-7)
    # public entry point: same result through window_adaptation.run
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    kernel = ab.nuts.new_kernel(srng, model)
    state, (step_size, imm_out), _ = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W)
    np.testing.assert_allclose(_np(step_size), ref["eps"], rtol=1e-7)
    np.testing.assert_allclose(_np(imm_out), ref["imm"], rtol=1e-7)
    np.testing.assert_allclose(_np(state.position), ref["q"], rtol=1e-6)


def test_window_adaptation_composed_matches_fused(ab):
    """window_adaptation(...) -> (init, update) built from the batched primitives gives the same warm-up."""
    rng = np.random.default_rng(21)
    C, W, d = 5, 24, 4
    mu, sigma = rng.standard_normal(d), np.exp(rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W, d)
    model = ab.models.IIDGaussian(mu, sigma)
    mk = lambda: ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    fused_state, (eps_f, imm_f), _ = ab.window_adaptation.run(ab.nuts.new_kernel(mk(), model),
                                                              ab.nuts.new_state(q0, model), W)
    # composed: one transition per call with per-transition slices of the injected draws
    init, update = ab.window_adaptation.window_adaptation(W)
    state = ab.nuts.new_state(q0, model)
    warm, params = init(state)
    for step in range(W):
        sl = {k: draws[k][:, step:step + 1] for k in ("z", "u_dir", "u_biased", "u_uniform")}
        kernel = ab.nuts.new_kernel(ab.InjectedDraws(**sl), model)
        info, _ = kernel(state, params[0], ab.metrics.per_chain(params[1]))
        warm, params = update(step, warm, params, info)
        s = info.state
        state = ab.integrators.IntegratorState(s.position, None, s.potential_energy, s.potential_energy_grad)
    np.testing.assert_allclose(_np(params[0]), _np(eps_f), rtol=1e-7)
    np.testing.assert_allclose(_np(params[1]), _np(imm_f), rtol=1e-7)
    np.testing.assert_allclose(_np(state.position), _np(fused_state.position), rtol=1e-6, atol=1e-9)


def gpu_hmc(ab, model, imm, q0, eps, draws, T, L, group=0):
    from aehmc_b200 import _engine
    srng = ab.InjectedDraws(z=draws["z"], u_accept=draws["u_accept"])
    info, extras = _engine.run("hmc", model, imm, srng, ab.hmc.new_state(q0, model), torch.as_tensor(eps, dtype=torch.float64),
                               n_transitions=T, num_integration_steps=L, store_draws=T, group=group)
    return dict(q=_np(info.state.position), p=_np(info.state.momentum), U=_np(info.state.potential_energy),
                g=_np(info.state.potential_energy_grad), acceptance_probability=_np(info.acceptance_probability),
                is_diverging=_np(info.is_diverging), draws=_np(extras["draws"]))


@pytest.mark.parametrize("d, group", [(6, 1), (6, 8), (100, 0)])
def test_hmc_iid_gaussian(ab, d, group):
    """Config 1 shape: HMC, velocity_verlet, L = 10, diagonal inverse mass matrix, iid Gaussian."""
    rng = np.random.default_rng(10 + d)
    C, T, L = 10, 4, 10
    mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    imm = sigma ** 2
    for eps in (0.25, 1.2):
        ref = parity.oracle_hmc(o_models.IIDGaussian(mu, sigma), q0, eps, imm, draws, T, L)
        got = gpu_hmc(ab, ab.models.IIDGaussian(mu, sigma), imm, q0, eps, draws, T, L, group)
        for k in ("q", "p", "g", "U", "acceptance_probability"):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-10, atol=1e-12, err_msg=k)
        np.testing.assert_array_equal(got["is_diverging"].astype(bool), ref["is_diverging"].astype(bool))
        np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-10, atol=1e-12)


def test_hmc_dense_metric_correlated_gaussian(ab):
    rng = np.random.default_rng(31)
    C, T, L, d = 8, 2, 7, 12
    mu, cov, prec = _corr_case(rng, d)
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_hmc(o_models.CorrelatedGaussian(mu, prec), q0, 0.2, cov, draws, T, L)
    got = gpu_hmc(ab, ab.models.CorrelatedGaussian(mu, prec), cov, q0, 0.2, draws, T, L)
    for k in ("q", "p", "g", "U", "acceptance_probability"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-11, err_msg=k)


@pytest.mark.parametrize("L", [1, 2])
def test_hmc_dense_metric_single_tick_transitions(ab, L):
    """A dense-metric transition that lasts ONE tick (HMC with L = 1) starts the next one on the following tick:
    the look-ahead momentum p0 AND its velocity imm.p0 must both have landed by then (they come from the same
    normals in one grouped launch, v0 = z.L^T)."""
    rng = np.random.default_rng(32 + L)
    C, T, d = 8, 5, 12
    mu, cov, prec = _corr_case(rng, d)
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_hmc(o_models.CorrelatedGaussian(mu, prec), q0, 0.2, cov, draws, T, L)
    got = gpu_hmc(ab, ab.models.CorrelatedGaussian(mu, prec), cov, q0, 0.2, draws, T, L)
    for k in ("q", "p", "g", "U", "acceptance_probability"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-11, err_msg=k)
    np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-9, atol=1e-11)


def test_nuts_dense_metric_first_step_divergence(ab):
    """NUTS with a dense metric where the very first leapfrog diverges (huge step size, low threshold): every
    transition is one tick long and still uses the momentum / velocity of its OWN transition."""
    from aehmc_b200 import _engine
    rng = np.random.default_rng(41)
    C, T, d = 6, 4, 10
    mu, cov, prec = _corr_case(rng, d)
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    eps = np.array([50.0, 0.2, 50.0, 0.3, 50.0, 50.0])
    ref = parity.oracle_nuts(o_models.CorrelatedGaussian(mu, prec), q0, eps, cov, draws, T, div_thr=10.0)
    assert ref["is_diverging"].astype(bool).sum() >= 3 and (ref["n_leapfrog"] == 1).sum() >= 3
    model = ab.models.CorrelatedGaussian(mu, prec)
    srng = ab.InjectedDraws(draws["z"], draws["u_dir"], draws["u_biased"], draws["u_uniform"])
    info, extras = _engine.run("nuts", model, cov, srng, ab.nuts.new_state(q0, model),
                               torch.as_tensor(eps, dtype=torch.float64), n_transitions=T, divergence_threshold=10.0,
                               store_draws=T)
    np.testing.assert_array_equal(_np(info.num_doublings), ref["num_doublings"])
    np.testing.assert_array_equal(_np(extras["n_leapfrog"]), ref["n_leapfrog"])
    np.testing.assert_array_equal(_np(info.is_diverging).astype(bool), ref["is_diverging"].astype(bool))
    np.testing.assert_allclose(_np(extras["draws"]), ref["draws"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(_np(info.state.momentum), ref["p"], rtol=1e-9, atol=1e-11)


def test_nuts_dense_metric_triangular_flag(ab, monkeypatch):
    """b2h_metric.reserved bit 0 (the factors are triangular: the momentum contractions of restarting chains skip the
    zero k-range of every column tile, gemm.cu) against the same run reading the factors in full; d = 300 spans three
    column tiles, so both triangles really skip."""
    from aehmc_b200 import _engine, _lib
    rng = np.random.default_rng(43)
    C, T, d = 6, 3, 300
    mu, cov, prec = _corr_case(rng, d)
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    model = ab.models.CorrelatedGaussian(mu, prec)

    def run():
        srng = ab.InjectedDraws(draws["z"], draws["u_dir"], draws["u_biased"], draws["u_uniform"])
        return _engine.run("nuts", model, cov, srng, ab.nuts.new_state(q0, model), 0.15, n_transitions=T, store_draws=T)

    info1, ex1 = run()
    flags = []
    full = ab.metrics.GaussianMetric.struct

    def struct_no_flag(self):
        m = full(self)
        flags.append(int(m.reserved))
        m.reserved = 0
        return m

    monkeypatch.setattr(ab.metrics.GaussianMetric, "struct", struct_no_flag)
    info0, ex0 = run()
    assert flags and all(f == 1 for f in flags)
    np.testing.assert_array_equal(_np(info1.num_doublings), _np(info0.num_doublings))
    np.testing.assert_array_equal(_np(ex1["n_leapfrog"]), _np(ex0["n_leapfrog"]))
    assert _np(ex1["n_leapfrog"]).sum() > 2 * C * T
    np.testing.assert_allclose(_np(ex1["draws"]), _np(ex0["draws"]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(_np(info1.state.momentum), _np(info0.state.momentum), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("metric_kind", ["diag", "dense"])
def test_nuts_gaussian_potential_formed_by_tick_kernel(ab, monkeypatch, metric_kind):
    """Correlated Gaussian target: the tile tick kernel forms U = 0.5 (q' - mu) . g' in its pass A (engine_split.inl,
    EngineView::u_center) -- against the oracle, and against the same run with the separate potential kernel
    (B2H_TICK_POTENTIAL=0).  d = 200: rows longer than one ring piece, i.e. the one-chain layouts that carry it."""
    from aehmc_b200 import _engine
    rng = np.random.default_rng(47)
    C, T, d = 5, 2, 200
    mu, cov, prec = _corr_case(rng, d)
    imm = cov if metric_kind == "dense" else np.diag(cov).copy()
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(o_models.CorrelatedGaussian(mu, prec), q0, 0.1, imm, draws, T)
    model = ab.models.CorrelatedGaussian(mu, prec)

    def run():
        srng = ab.InjectedDraws(draws["z"], draws["u_dir"], draws["u_biased"], draws["u_uniform"])
        return _engine.run("nuts", model, imm, srng, ab.nuts.new_state(q0, model), 0.1, n_transitions=T, store_draws=T)

    info1, ex1 = run()
    monkeypatch.setenv("B2H_TICK_POTENTIAL", "0")
    info0, ex0 = run()
    for info, ex in ((info1, ex1), (info0, ex0)):
        np.testing.assert_array_equal(_np(info.num_doublings), ref["num_doublings"])
        np.testing.assert_array_equal(_np(ex["n_leapfrog"]), ref["n_leapfrog"])
        np.testing.assert_allclose(_np(ex["draws"]), ref["draws"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(_np(info.state.potential_energy), ref["U"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(_np(info1.acceptance_probability), _np(info0.acceptance_probability), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("metric_kind", ["diag", "dense"])
def test_hmc_proposal_closure_matches_oracle(ab, metric_kind):
    """hmc.hmc_proposal(integrator, kinetic_energy, L, threshold) -> propose(srng, state, step_size)
    (reference hmc.py:129-206), composed the way hmc.new_kernel.step composes it (hmc.py:110-123)."""
    rng = np.random.default_rng(51)
    C, T, L, d = 9, 3, 6, 7
    mu, cov, prec = _corr_case(rng, d)
    imm = cov if metric_kind == "dense" else np.diag(cov).copy()
    q0 = mu + rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_hmc(o_models.CorrelatedGaussian(mu, prec), q0, 0.3, imm, draws, T, L, div_thr=0.05)
    model = ab.models.CorrelatedGaussian(mu, prec)
    momentum_generator, kinetic_energy, _ = ab.metrics.gaussian_metric(imm)
    integrator = ab.integrators.velocity_verlet(model, kinetic_energy)
    propose = ab.hmc.hmc_proposal(integrator, kinetic_energy, L, 0.05)
    srng = ab.InjectedDraws(z=draws["z"], u_accept=draws["u_accept"])
    state = ab.hmc.new_state(q0, model)
    for t in range(T):
        state = state._replace(momentum=momentum_generator(srng, num_chains=C, dim=d, transition=t))
        info, updates = propose(srng, state, 0.3)
        np.testing.assert_allclose(_np(info.state.position), ref["draws"][t], rtol=1e-9, atol=1e-11)
        state = info.state
    assert info.num_doublings is None and info.is_turning is None and updates == {}
    np.testing.assert_allclose(_np(info.state.momentum), ref["p"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(_np(info.state.potential_energy), ref["U"], rtol=1e-9)
    np.testing.assert_allclose(_np(info.state.potential_energy_grad), ref["g"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(_np(info.acceptance_probability), ref["acceptance_probability"], rtol=1e-9)
    np.testing.assert_array_equal(_np(info.is_diverging).astype(bool), ref["is_diverging"].astype(bool))
    assert ref["is_diverging"].any() and not ref["is_diverging"].all()
    with pytest.raises(ValueError):
        propose(srng, ab.hmc.new_state(q0, model), 0.3)


def test_hmc_public_kernel_signature(ab):
    """hmc.new_kernel(srng, logprob_fn)(state, step_size, imm, L) -> (Diagnostics, updates)."""
    model = ab.models.IIDGaussian(np.zeros(3), np.ones(3))
    kernel = ab.hmc.new_kernel(ab.RandomStream(seed=1), model)
    state = ab.hmc.new_state(np.zeros((16, 3)), model)
    info, updates = kernel(state, 0.3, np.ones(3), 10)
    assert info.state.position.shape == (16, 3) and info.num_doublings is None and info.is_turning is None
    assert info.acceptance_probability.shape == (16,) and updates["n_leapfrog"].tolist() == [10] * 16


def test_native_rng_replays_through_injected_draws(ab):
    """Native (Philox) mode and validation mode fed with b2h_philox_fill's export take identical decisions,
    so a native run can be replayed by the oracle."""
    import ctypes as C
    from aehmc_b200 import _engine, _lib, backend
    rng = np.random.default_rng(5)
    Cn, T, d, maxd = 32, 3, 10, 10
    model = ab.models.NealFunnel(d)
    q0 = rng.standard_normal((Cn, d))
    native, ex = _engine.run("nuts", model, np.ones(d), ab.RandomStream(seed=123), ab.nuts.new_state(q0, model), 0.3,
                             n_transitions=T, store_draws=T)
    dev = model.device
    z = torch.empty((Cn, T, d), dtype=torch.float64, device=dev)
    ud = torch.empty((Cn, T, maxd), dtype=torch.float64, device=dev); ub = torch.empty_like(ud)
    uu = torch.empty((Cn, T, 1023), dtype=torch.float64, device=dev)
    _lib.check(_lib.load().b2h_philox_fill(backend.context(dev), C.c_uint64(123), C.c_uint64(0), C.c_uint64(0),
                                           C.c_int64(Cn), C.c_int64(T), C.c_int64(d), C.c_int32(maxd), backend.ptr(z),
                                           backend.ptr(ud), backend.ptr(ub), backend.ptr(uu), None))
    replay, ex2 = _engine.run("nuts", model, np.ones(d), ab.InjectedDraws(z, ud, ub, uu), ab.nuts.new_state(q0, model),
                              0.3, n_transitions=T, store_draws=T)
    assert torch.equal(ex["draws"], ex2["draws"]) and torch.equal(native.num_doublings, replay.num_doublings)
    draws = {"z": _np(z), "u_dir": _np(ud), "u_biased": _np(ub), "u_uniform": _np(uu), "u_accept": np.zeros((Cn, T))}
    ref = parity.oracle_nuts(o_models.NealFunnel(d), q0, 0.3, np.ones(d), draws, T)
    np.testing.assert_array_equal(_np(native.num_doublings), ref["num_doublings"])
    np.testing.assert_allclose(_np(native.state.position), ref["q"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("which", ["logistic", "user"])
def test_window_adaptation_in_split_mode(ab, which):
    """Warm-up in the per-tick (split) engine: logistic regression (FMA path) and a user-written model go through the
    same dual-averaging / Welford schedule as the oracle under injected draws (short run: adaptation is a feedback
    loop)."""
    rng = np.random.default_rng(23)
    C, W = 6, 22
    if which == "logistic":
        N, d = 96, 4
        X = rng.standard_normal((N, d))
        y = (rng.random(N) < 0.5).astype(np.float64)
        o_model = o_models.LogisticRegression(X, y, 1.0)
        model = ab.models.LogisticRegression(X, y, 1.0)
    else:
        d = 4
        src = r"""
        template <typename S, typename T>
        __device__ S log_density(const S* q, int d, const T* data) {
            S lp = (T)0;
            for (int i = 0; i < d; ++i) lp -= (T)0.5 * square(q[i] - data[i]) * data[d + i];
            return lp;
        }"""
        mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
        o_model = o_models.IIDGaussian(mu, sigma)
        model = ab.models.UserModel(src, d, data=np.concatenate([mu, 1.0 / sigma ** 2]), autodiff=True)
    q0 = 0.5 * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W, d)
    ref = parity.oracle_nuts(o_model, q0, 1.0, np.ones(d), draws, W, schedule_steps=W)
    srng = ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform")})
    kernel = ab.nuts.new_kernel(srng, model)
    state, (step_size, imm_out), _ = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), W)
    np.testing.assert_allclose(_np(step_size), ref["eps"], rtol=1e-6)
    np.testing.assert_allclose(_np(imm_out), ref["imm"], rtol=1e-6)
    np.testing.assert_allclose(_np(state.position), ref["q"], rtol=1e-5, atol=1e-8)


def test_aesara_op_shim_perform_marshals_to_the_kernels(ab, monkeypatch):
    """Aesara is not installable here, but `perform` (the only logic-free marshalling of aesara_ops.py) can be driven
    directly: NumPy in, NumPy out, identical to the kernel called through the Python API."""
    from aehmc_b200 import aesara_ops
    monkeypatch.setattr(aesara_ops, "HAVE_AESARA", True)
    rng = np.random.default_rng(41)
    C, d = 9, 4
    mu, sigma = rng.standard_normal(d), np.exp(0.3 * rng.standard_normal(d))
    model = ab.models.IIDGaussian(mu, sigma)
    q0 = rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, 1, d)
    mk = lambda: ab.InjectedDraws(**{k: draws[k] for k in ("z", "u_dir", "u_biased", "u_uniform", "u_accept")})
    U0, g0 = model.potential_and_grad(q0)
    pot = aesara_ops.PotentialAndGradOp(model)
    store = [[None], [None]]
    pot.perform(None, [q0], store)
    np.testing.assert_array_equal(store[0][0], _np(U0))
    np.testing.assert_array_equal(store[1][0], _np(g0))

    op = aesara_ops.NUTSStepOp(mk(), model)
    out = [[None] for _ in range(8)]
    op.perform(None, [q0, _np(U0), _np(g0), np.full(C, 0.3), sigma ** 2], out)
    info, _ = ab.nuts.new_kernel(mk(), model)(ab.nuts.new_state(q0, model), 0.3, sigma ** 2)
    np.testing.assert_array_equal(out[0][0], _np(info.state.position))
    np.testing.assert_array_equal(out[4][0], _np(info.acceptance_probability))
    np.testing.assert_array_equal(out[5][0], _np(info.num_doublings))

    hop = aesara_ops.HMCStepOp(mk(), model, num_integration_steps=5)
    hout = [[None] for _ in range(6)]
    hop.perform(None, [q0, _np(U0), _np(g0), np.full(C, 0.2), sigma ** 2], hout)
    hinfo, _ = ab.hmc.new_kernel(mk(), model)(ab.hmc.new_state(q0, model), 0.2, sigma ** 2, 5)
    np.testing.assert_array_equal(hout[0][0], _np(hinfo.state.position))
    np.testing.assert_array_equal(hout[5][0], _np(hinfo.is_diverging))
