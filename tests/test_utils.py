"""RaveledParamsMap: mirrors the reference's tests/test_utils.py (the same beta / tau / kappa layout and the dtype
round trip), plus the chains axis this build adds."""
import numpy as np
import pytest

from aehmc_b200.utils import ParamSpec, RaveledParamsMap


def _layout():
    beta = ParamSpec("beta", (3, 2))
    tau = ParamSpec("tau", ())
    kappa = ParamSpec("kappa", (20,))
    return beta, tau, kappa


def test_RaveledParamsMap():
    # reference tests/test_utils.py:11-57
    beta, tau, kappa = _layout()
    rp_map = RaveledParamsMap([beta, tau, kappa])
    assert repr(rp_map) == "RaveledParamsMap((beta, tau, kappa))"
    exp_beta = np.exp(np.arange(6)).reshape(3, 2)
    exp_tau = 1.0
    exp_kappa = np.exp(np.arange(20))
    expected = np.concatenate([exp_beta.ravel(), np.atleast_1d(exp_tau), exp_kappa.ravel()])
    raveled = rp_map.ravel_params([exp_beta, exp_tau, exp_kappa])
    assert np.array_equal(raveled, expected)
    parts = rp_map.unravel_params(expected)
    assert np.array_equal(parts[beta], exp_beta)
    assert np.array_equal(parts[tau], exp_tau) and parts[tau].shape == ()
    assert np.array_equal(parts[kappa], exp_kappa)


def test_RaveledParamsMap_dtype():
    # reference tests/test_utils.py:60-79: every part comes back in its reference dtype
    tau = ParamSpec("tau", (), np.float64)
    lmbda = ParamSpec("lmbda", (), np.int64)
    rp_map = RaveledParamsMap([tau, lmbda])
    q = rp_map.ravel_params((0.25, 7))
    parts = rp_map.unravel_params(q)
    assert parts[tau].dtype == np.float64 and parts[lmbda].dtype == np.int64
    assert parts[lmbda] == 7


def test_RaveledParamsMap_templates_and_chains():
    # templates instead of specs; a leading chains axis maps to the leading axis of q
    rng = np.random.default_rng(0)
    mu = np.zeros(())
    theta = np.zeros((8,), dtype=np.float32)
    rp_map = RaveledParamsMap([mu, theta])
    assert rp_map.size == 9
    C = 5
    mu_c, theta_c = rng.standard_normal(C), rng.standard_normal((C, 8)).astype(np.float32)
    q = rp_map.ravel_params([mu_c, theta_c])
    assert q.shape == (C, 9)
    parts = rp_map.unravel_params(q)
    vals = list(parts.values())
    assert np.array_equal(vals[0], mu_c) and vals[1].dtype == np.float32 and np.array_equal(vals[1], theta_c)
    # an unbatched parameter is broadcast over the chains of the others
    q2 = rp_map.ravel_params([0.5, theta_c])
    assert np.array_equal(q2[:, 0], np.full(C, 0.5))
    with pytest.raises(ValueError):
        rp_map.ravel_params([mu_c, theta_c[:3]])
    with pytest.raises(ValueError):
        rp_map.unravel_params(np.zeros(7))


def test_RaveledParamsMap_torch():
    torch = pytest.importorskip("torch")
    beta, tau, kappa = _layout()
    rp_map = RaveledParamsMap([beta, tau, kappa])
    C = 3
    b, t, k = torch.randn(C, 3, 2, dtype=torch.float64), torch.randn(C, dtype=torch.float64), torch.randn(C, 20, dtype=torch.float64)
    q = rp_map.ravel_params([b, t, k])
    assert isinstance(q, torch.Tensor) and q.shape == (C, 27)
    parts = rp_map.unravel_params(q)
    assert torch.equal(parts[beta], b) and torch.equal(parts[tau], t) and torch.equal(parts[kappa], k)
