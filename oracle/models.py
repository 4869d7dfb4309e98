"""Log-density models for the oracle (test infrastructure).

The reference takes a user ``logprob_fn`` and differentiates it with
``aesara.grad`` (reference hmc.py:33-34, integrators.py:64-65).  Here a model is
an object with ``potential_and_grad(q) -> (U, dU/dq)`` where ``U = -logp``.
``Normal`` restates the aeppl normal log-density the reference's README and
tests use; the others are the built-in targets of the CUDA engine
(SURVEY.md section 8a), defined here first so the CUDA device functions have a
CPU statement to be checked against.
"""
from __future__ import annotations

import numpy as np

_LOG_SQRT_2PI = float(np.log(np.sqrt(2.0 * np.pi)))


class Normal:
    """N(mu, sigma) on a 0-d or 1-d position, constants kept (aeppl logprob)."""

    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)

    def potential_and_grad(self, q):
        r = (q - self.mu) / self.sigma
        U = np.sum(0.5 * r * r + _LOG_SQRT_2PI + np.log(self.sigma))
        return U, r / self.sigma


class IIDGaussian:
    """U = 1/2 sum r_i g_i + const, r = q - mu, g = r * inv_var (inv_var = 1/sigma^2)."""

    def __init__(self, mu, sigma, const=0.0):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.sigma = np.asarray(sigma, dtype=np.float64)
        self.inv_var = 1.0 / (self.sigma * self.sigma)
        self.const = float(const)

    def potential_and_grad(self, q):
        r = q - self.mu
        g = r * self.inv_var
        return 0.5 * np.sum(r * g) + self.const, g


class CorrelatedGaussian:
    """U = 1/2 r^T Lambda r with a symmetric precision matrix Lambda."""

    def __init__(self, mu, precision):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.precision = np.asarray(precision, dtype=np.float64)

    def potential_and_grad(self, q):
        r = q - self.mu
        g = self.precision @ r
        return 0.5 * np.dot(r, g), g


class NealFunnel:
    """q = (v, x_1..x_{d-1}); v ~ N(0, 3), x_i ~ N(0, exp(v/2)); constants dropped."""

    def __init__(self, dim=10):
        self.dim = int(dim)

    def potential_and_grad(self, q):
        v, x = q[0], q[1:]
        n = self.dim - 1
        ev = np.exp(-v)
        ss = np.sum(x * x)
        U = v * v / 18.0 + 0.5 * ev * ss + 0.5 * n * v
        g = np.empty_like(q)
        g[0] = v / 9.0 - 0.5 * ev * ss + 0.5 * n
        g[1:] = x * ev
        return U, g


EIGHT_SCHOOLS_Y = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
EIGHT_SCHOOLS_SIGMA = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])


class EightSchools:
    """Non-centred eight schools, q = (mu, log tau, theta~_1..J).

    mu ~ N(0,5); tau ~ HalfCauchy(5) sampled on log tau (Jacobian included);
    theta~ ~ N(0,1); y_j ~ N(mu + tau theta~_j, sigma_j).  Constants dropped.
    """

    def __init__(self, y=EIGHT_SCHOOLS_Y, sigma=EIGHT_SCHOOLS_SIGMA):
        self.y = np.asarray(y, dtype=np.float64)
        self.sigma = np.asarray(sigma, dtype=np.float64)
        self.inv_var = 1.0 / (self.sigma * self.sigma)
        self.dim = 2 + self.y.shape[0]

    def potential_and_grad(self, q):
        mu, t, th = q[0], q[1], q[2:]
        tau = np.exp(t)
        a = tau * tau / 25.0
        resid = self.y - mu - tau * th
        w = resid * self.inv_var
        U = mu * mu / 50.0 - t + np.log1p(a) + 0.5 * np.sum(th * th) + 0.5 * np.sum(resid * w)
        g = np.empty_like(q)
        g[0] = mu / 25.0 - np.sum(w)
        g[1] = -1.0 + 2.0 * a / (1.0 + a) - tau * np.sum(w * th)
        g[2:] = th - tau * w
        return U, g


class LogisticRegression:
    """U = sum softplus(s_i) - y_i s_i + 1/2 |b|^2 / prior_scale^2, s = X b."""

    def __init__(self, X, y, prior_scale=1.0):
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.inv_prior_var = 1.0 / float(prior_scale) ** 2

    def potential_and_grad(self, q):
        s = self.X @ q
        softplus = np.maximum(s, 0.0) + np.log1p(np.exp(-np.abs(s)))
        U = np.sum(softplus - self.y * s) + 0.5 * self.inv_prior_var * np.dot(q, q)
        sig = 0.5 * (1.0 + np.tanh(0.5 * s))
        g = self.X.T @ (sig - self.y) + self.inv_prior_var * q
        return U, g


class Callable:
    """Wrap a plain ``f(q) -> (U, g)`` (used for the reference's physics toys)."""

    def __init__(self, fn):
        self.fn = fn

    def potential_and_grad(self, q):
        return self.fn(q)
