"""Native-RNG (Philox) distributional checks on known posteriors: the statistical tests of the reference
(tests/test_hmc.py:158-346 MCSE tests, :13-97 warm-up ranges, tests/test_step_size.py) at many-chain scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ab(cuda_device):
    import aehmc_b200
    return aehmc_b200


def test_nuts_correlated_gaussian_moments(ab):
    """reference tests/test_hmc.py:296-346 target (2-d, rho = 0.5, sigma = (1, 2)), 4096 chains x 60 draws in
    native-RNG mode.  The reference algorithm (2**k + 1 leapfrogs per sub-tree) is not exactly invariant, so
    the comparison is with the long-run moments of the ORACLE chain (tests/golden/nuts_moments.json), which
    the GPU path must share; the analytic posterior is only checked loosely."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "nuts_moments.json")))
    loc = np.array(gold["target"]["loc"]); scale = np.array(gold["target"]["scale"]); rho = gold["target"]["rho"]
    cov = np.array([[scale[0] ** 2, rho * scale[0] * scale[1]], [rho * scale[0] * scale[1], scale[1] ** 2]])
    model = ab.models.CorrelatedGaussian(loc, np.linalg.inv(cov))
    C = 4096
    q0 = np.tile(np.array([[1.0, 1.0]]), (C, 1))
    for case in gold["cases"]:
        kernel = ab.nuts.new_kernel(ab.RandomStream(seed=0), model)
        info, draws, stats, _ = ab.sampling.sample(kernel, ab.nuts.new_state(q0, model), case["step_size"],
                                                   np.array(case["inverse_mass_matrix"]), 60)
        x = draws[20:].reshape(-1, 2).double().cpu().numpy()
        print("nuts moments", case["step_size"], x.mean(0), x.var(0), np.corrcoef(x.T)[0, 1], "oracle", case)
        np.testing.assert_allclose(x.mean(0), case["mean"], atol=0.06)
        np.testing.assert_allclose(x.var(0), case["var"], rtol=0.05)
        assert abs(np.corrcoef(x.T)[0, 1] - case["corr"]) < 0.03
        assert np.all(np.abs(x.var(0) / scale ** 2 - 1) < 0.45)       # loosely the posterior
        assert (stats[..., 3].long() & 2).sum().item() == 0            # no divergence flagged (bit 1 of the flags)


def test_hmc_iid_gaussian_moments_and_ks(ab):
    from scipy import stats as sstats
    d, C = 4, 8192
    mu = np.array([1.0, -2.0, 0.5, 3.0]); sigma = np.array([1.0, 2.0, 0.5, 1.5])
    model = ab.models.IIDGaussian(mu, sigma)
    kernel = ab.hmc.new_kernel(ab.RandomStream(seed=3), model)
    # eps * L = 3.5 rad of the (preconditioned) oscillator: far from the 2 pi resonance
    info, draws, _, _ = ab.sampling.sample(kernel, ab.hmc.new_state(np.zeros((C, d)), model), 0.7, sigma ** 2, 30,
                                           num_integration_steps=5)
    x = draws[-1].double().cpu().numpy()               # one draw per chain: independent samples
    for j in range(d):
        assert sstats.kstest((x[:, j] - mu[j]) / sigma[j], "norm").pvalue > 1e-3
    assert info.acceptance_probability.mean().item() > 0.6


def test_hmc_stability_limit(ab):
    """reference tests/test_hmc.py:100-155: N(1, 2), identity metric, L = 30: step 3.9 samples, 4.1 is stuck."""
    model = ab.models.IIDGaussian([1.0], [2.0])
    C = 2048
    for eps, ok in ((3.9, True), (4.1, False)):
        kernel = ab.hmc.new_kernel(ab.RandomStream(seed=0), model)
        info, draws, _, _ = ab.sampling.sample(kernel, ab.hmc.new_state(np.full((C, 1), 3.0), model), eps, 1.0, 200,
                                               num_integration_steps=30)
        x = draws[100:].double().cpu().numpy()
        if ok:
            assert abs(x.mean() - 1.0) < 0.1 and abs(x.var() - 4.0) < 0.4
        else:
            assert np.all(x == 3.0)


def test_window_adaptation_reaches_target(ab):
    """reference tests/test_hmc.py:13-52 (N(1,2) warm-up): 0.1 < step size < 2, imm ~ sigma^2 within 100 %;
    plus mean acceptance near the 0.8 target afterwards."""
    model = ab.models.IIDGaussian([1.0, -1.0, 0.0], [2.0, 0.5, 1.0])
    C = 1024
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=0), model)
    state, (eps, imm), _ = ab.window_adaptation.run(kernel, ab.nuts.new_state(np.full((C, 3), 3.0), model), 400)
    eps, imm = eps.cpu().numpy(), imm.cpu().numpy()
    assert np.all(eps > 0.1) and np.all(eps < 2.5)
    assert np.all(np.abs(np.median(imm, axis=0) / np.array([4.0, 0.25, 1.0]) - 1) < 0.5)
    info, draws, stats, _ = ab.sampling.sample(kernel, state, torch.as_tensor(eps), ab.metrics.per_chain(torch.as_tensor(imm)), 50)
    # the reference's dual averaging (shrinkage point = the step size itself, averaged OLD iterates; SURVEY
    # Q16) and its acceptance statistic (last sub-tree only, Q9) settle near, not at, the 0.8 target
    assert abs(stats[..., 0].mean().item() - 0.8) < 0.15


def test_funnel_divergences_and_depth_spread(ab):
    model = ab.models.NealFunnel(10)
    C = 4096
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=5), model)
    q0 = np.random.default_rng(0).standard_normal((C, 10))
    info, draws, stats, _ = ab.sampling.sample(kernel, ab.nuts.new_state(q0, model), 0.2, np.ones(10), 40)
    depth = stats[..., 1].cpu().numpy()
    assert depth.min() >= 1 and depth.max() <= 10 and len(np.unique(depth)) >= 4
    v = draws[10:, :, 0].double().cpu().numpy()
    assert abs(v.mean()) < 1.0 and 1.0 < v.std() < 4.0


@pytest.mark.parametrize("engine", ["fused", "split"])
def test_nuts_exact_doubling_matches_the_posterior(ab, engine):
    """The non-reference option exact_doubling=True (balanced sub-trees): on the reference's MCSE target
    (tests/test_hmc.py:170-187: 2-d Gaussian, sigma = (1, 2), rho = 0.5) the moments match the ANALYTIC posterior
    within Monte-Carlo error in both engines, where the reference behaviour is off by 7-34 % in variance
    (DESIGN.md 2.1); and at eps = 0.5 on logprob = -2 (x - 1)^2 the variance is 1/4, not the reference's 0.002."""
    scale, rho = np.array([1.0, 2.0]), 0.5
    cov = np.array([[1.0, rho * 2.0], [rho * 2.0, 4.0]])
    C = 4096
    if engine == "fused":      # diagonal metric, elementwise model: independent coordinates with the same marginals
        model, imm, target_corr = ab.models.IIDGaussian(np.zeros(2), scale), np.ones(2), 0.0
    else:                      # dense metric: split engine
        model, imm, target_corr = ab.models.CorrelatedGaussian(np.zeros(2), np.linalg.inv(cov)), np.eye(2), rho
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=1), model, exact_doubling=True)
    info, draws, stats, _ = ab.sampling.sample(kernel, ab.nuts.new_state(np.ones((C, 2)), model), 0.5, imm, 70)
    x = draws[20:].reshape(-1, 2).double().cpu().numpy()
    np.testing.assert_allclose(x.mean(0), 0.0, atol=0.04)
    np.testing.assert_allclose(x.var(0), scale ** 2, rtol=0.04)
    assert abs(np.corrcoef(x.T)[0, 1] - target_corr) < 0.02

    quad = ab.models.IIDGaussian([1.0], [0.5])                         # logprob = -2 (x - 1)^2 + const
    kernel = ab.nuts.new_kernel(ab.RandomStream(seed=2), quad, exact_doubling=True)
    info, draws, _, _ = ab.sampling.sample(kernel, ab.nuts.new_state(np.zeros((C, 1)), quad), 0.5, np.ones(1), 60)
    v = draws[20:].double().cpu().numpy().var()
    assert abs(v - 0.25) < 0.02
