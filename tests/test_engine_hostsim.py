"""engine.cuh / models.cuh (the CUDA chain state machine and model gradients) compiled with
g++ through tests/host_sim and compared with the oracle under injected draws.  This checks the
device code's LOGIC where there is no GPU; the same comparisons run against the real kernels
in tests/test_gpu_*.py."""
import shutil
import sys
import os

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "host_sim"))
import hostsim  # noqa: E402
import parity  # noqa: E402
from oracle import adaptation, kernels, models, streams  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


def test_readme_golden_bit_exact_through_engine_code():
    rec = streams.RecordingStreamDraws(0, "nuts")
    model = models.IIDGaussian([0.0], [1.0], const=models._LOG_SQRT_2PI)
    kernel = kernels.nuts_new_kernel(rec, model)
    info, extras = kernel(kernels.new_state(np.zeros(1), model), 1e-2, np.float64(1.0))
    assert float(info.state.position[0]) == 1.1034719409361107
    inj = {k: v[None] for k, v in rec.injected().items()}
    out = hostsim.run(0, [0.0], [1.0], models._LOG_SQRT_2PI, 1.0, np.zeros((1, 1)), 1e-2, inj, 1)
    assert float(out["q"][0, 0]) == 1.1034719409361107           # reference README.md:54
    assert out["num_doublings"][0] == 8 and out["n_leapfrog"][0] == 136
    assert out["acceptance_probability"][0] == pytest.approx(info.acceptance_probability, rel=1e-14)


@pytest.mark.parametrize("d, imm_kind, eps", [(1, "scalar", 0.3), (5, "diag", 0.4), (5, "diag", 1.7),
                                             (3, "per_chain", 0.2), (7, "scalar", 1e-3), (2, "diag", 30.0)])
def test_nuts_iid_gaussian(d, imm_kind, eps):
    rng = np.random.default_rng(100 + d)
    C, T = 12, 3
    mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
    imm = {"scalar": np.float64(0.7), "diag": sigma ** 2 * np.exp(0.2 * rng.standard_normal(d)),
           "per_chain": np.exp(0.3 * rng.standard_normal((C, d)))}[imm_kind]
    q0 = mu + sigma * rng.standard_normal((C, d))
    eps_c = eps * np.exp(0.3 * rng.standard_normal(C))
    maxd = 6 if eps < 0.01 else 10
    draws = parity.random_draws(rng, C, T, d, maxd)
    model = models.IIDGaussian(mu, sigma, const=0.25)
    ref = parity.oracle_nuts(model, q0, eps_c, imm, draws, T, maxd=maxd, per_chain_imm=imm_kind == "per_chain")
    got = hostsim.run(0, mu, model.inv_var, 0.25, imm, q0, eps_c, draws, T, maxd=maxd, n_store=T)
    parity.assert_nuts_parity(got, ref, rtol=1e-11, what=f"iid d={d}")
    np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("eps", [0.05, 0.5, 3.0])
def test_nuts_funnel(eps):
    rng = np.random.default_rng(7)
    C, T, d = 16, 3, 10
    q0 = rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    model = models.NealFunnel(d)
    ref = parity.oracle_nuts(model, q0, eps, np.ones(d), draws, T)
    got = hostsim.run(2, None, None, 0.0, np.ones(d), q0, eps, draws, T, n_store=T)
    parity.assert_nuts_parity(got, ref, rtol=1e-10, what="funnel")
    assert ref["is_diverging"].any() or eps < 1.0


def test_nuts_eight_schools():
    rng = np.random.default_rng(8)
    C, T, d = 16, 4, 10
    q0 = 0.5 * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    model = models.EightSchools()
    ref = parity.oracle_nuts(model, q0, 0.3, np.ones(d), draws, T)
    got = hostsim.run(3, model.y, model.inv_var, 0.0, np.ones(d), q0, 0.3, draws, T, n_store=T)
    parity.assert_nuts_parity(got, ref, rtol=1e-10, what="eight schools")
    assert len(set(ref["num_doublings"].tolist())) > 1          # heterogeneous tree depths


def test_window_adaptation_long_schedule():
    """window_adaptation.run (reference window_adaptation.py:17-116) per chain over a 200-step
    schedule: dual averaging + Welford + two slow-window ends + final averaged step size, then 2
    draws.  d = 1 keeps every reduction a single term, so oracle and engine perform the same IEEE
    operations and the comparison stays tight over 200 transitions (with d > 1 the adaptation
    feedback amplifies summation-order rounding ~2.5x per transition)."""
    rng = np.random.default_rng(9)
    C, W, extra, d = 6, 200, 2, 1
    mu, sigma = np.array([1.0]), np.array([2.0])
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W + extra, d)
    model = models.IIDGaussian(mu, sigma)
    sched = adaptation.build_schedule(W)
    assert sum(e for _, e in sched) == 2
    ref = parity.oracle_nuts(model, q0, 1.0, np.ones(d), draws, W + extra, schedule_steps=W)
    got = hostsim.run(0, mu, model.inv_var, 0.0, np.ones((C, d)), q0, 1.0, draws, W + extra, schedule=sched,
                      n_store=W + extra)
    parity.assert_nuts_parity(got, ref, rtol=1e-12, what="adapt")
    np.testing.assert_allclose(got["eps"], ref["eps"], rtol=1e-12)
    np.testing.assert_allclose(got["imm"], ref["imm"], rtol=1e-12)
    np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-12, atol=1e-14)
    assert np.all(ref["imm"] > 0.5) and np.all(ref["imm"] < 12.0)       # adapted towards sigma^2 = 4


def test_window_adaptation_gaussian_short():
    rng = np.random.default_rng(19)
    C, W, d = 4, 22, 5
    mu, sigma = rng.standard_normal(d), np.exp(rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W, d)
    model = models.IIDGaussian(mu, sigma)
    sched = adaptation.build_schedule(W)
    assert sum(e for _, e in sched) == 1
    ref = parity.oracle_nuts(model, q0, 1.0, np.ones(d), draws, W, schedule_steps=W)
    got = hostsim.run(0, mu, model.inv_var, 0.0, np.ones((C, d)), q0, 1.0, draws, W, schedule=sched)
    parity.assert_nuts_parity(got, ref, rtol=1e-7, what="adapt")
    np.testing.assert_allclose(got["eps"], ref["eps"], rtol=1e-8)
    np.testing.assert_allclose(got["imm"], ref["imm"], rtol=1e-8)


def test_window_adaptation_funnel_short():
    """Same on a chaotic target, kept short (25 steps, one slow-window end at step 21)."""
    rng = np.random.default_rng(9)
    C, W, d = 6, 25, 10
    q0 = rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, W, d)
    model = models.NealFunnel(d)
    sched = adaptation.build_schedule(W)
    ref = parity.oracle_nuts(model, q0, 1.0, np.ones(d), draws, W, schedule_steps=W)
    got = hostsim.run(2, None, None, 0.0, np.ones((C, d)), q0, 1.0, draws, W, schedule=sched)
    parity.assert_nuts_parity(got, ref, rtol=1e-6, what="adapt funnel")
    np.testing.assert_allclose(got["eps"], ref["eps"], rtol=1e-7)
    np.testing.assert_allclose(got["imm"], ref["imm"], rtol=1e-7)


def test_hmc_iid_gaussian():
    rng = np.random.default_rng(10)
    C, T, d, L = 10, 4, 6, 10
    mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    draws = parity.random_draws(rng, C, T, d)
    model = models.IIDGaussian(mu, sigma)
    imm = sigma ** 2
    for eps in (0.25, 1.2):
        ref = parity.oracle_hmc(model, q0, eps, imm, draws, T, L)
        got = hostsim.run(0, mu, model.inv_var, 0.0, imm, q0, eps, draws, T, hmc_L=L, n_store=T)
        for k in ("q", "p", "g", "U", "acceptance_probability"):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-11, atol=1e-13, err_msg=k)
        np.testing.assert_array_equal(got["is_diverging"].astype(bool), ref["is_diverging"])
        np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-11, atol=1e-13)


# ---- the same state machine with the integration front held in "registers" (engine.cuh RegFront) -------------
@pytest.mark.parametrize("case", ["iid", "funnel", "schools", "adapt", "hmc"])
def test_register_front_matches_oracle(case):
    rng = np.random.default_rng(123)
    if case == "iid":
        C, T, d = 10, 3, 7
        mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
        q0 = mu + sigma * rng.standard_normal((C, d))
        eps = 0.5 * np.exp(0.3 * rng.standard_normal(C))
        draws = parity.random_draws(rng, C, T, d)
        model = models.IIDGaussian(mu, sigma, const=0.1)
        ref = parity.oracle_nuts(model, q0, eps, sigma ** 2, draws, T)
        got = hostsim.run(0, mu, model.inv_var, 0.1, sigma ** 2, q0, eps, draws, T, n_store=T, reg_front=True)
        parity.assert_nuts_parity(got, ref, rtol=1e-11, what="iid regs")
        np.testing.assert_allclose(got["draws"], ref["draws"], rtol=1e-11, atol=1e-13)
    elif case == "funnel":
        C, T, d = 16, 3, 10
        q0 = rng.standard_normal((C, d))
        draws = parity.random_draws(rng, C, T, d)
        for eps in (0.1, 3.0):
            ref = parity.oracle_nuts(models.NealFunnel(d), q0, eps, np.ones(d), draws, T)
            got = hostsim.run(2, None, None, 0.0, np.ones(d), q0, eps, draws, T, reg_front=True)
            parity.assert_nuts_parity(got, ref, rtol=1e-10, what="funnel regs")
    elif case == "schools":
        C, T, d = 16, 3, 10
        q0 = 0.5 * rng.standard_normal((C, d))
        draws = parity.random_draws(rng, C, T, d)
        model = models.EightSchools()
        ref = parity.oracle_nuts(model, q0, 0.3, np.ones(d), draws, T)
        got = hostsim.run(3, model.y, model.inv_var, 0.0, np.ones(d), q0, 0.3, draws, T, reg_front=True)
        parity.assert_nuts_parity(got, ref, rtol=1e-10, what="schools regs")
    elif case == "adapt":
        C, W, d = 4, 22, 5
        mu, sigma = rng.standard_normal(d), np.exp(rng.standard_normal(d))
        q0 = mu + sigma * rng.standard_normal((C, d))
        draws = parity.random_draws(rng, C, W, d)
        model = models.IIDGaussian(mu, sigma)
        sched = adaptation.build_schedule(W)
        ref = parity.oracle_nuts(model, q0, 1.0, np.ones(d), draws, W, schedule_steps=W)
        got = hostsim.run(0, mu, model.inv_var, 0.0, np.ones((C, d)), q0, 1.0, draws, W, schedule=sched, reg_front=True)
        parity.assert_nuts_parity(got, ref, rtol=1e-7, what="adapt regs")
        np.testing.assert_allclose(got["eps"], ref["eps"], rtol=1e-8)
    else:
        C, T, d, L = 8, 3, 6, 10
        mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
        q0 = mu + sigma * rng.standard_normal((C, d))
        draws = parity.random_draws(rng, C, T, d)
        model = models.IIDGaussian(mu, sigma)
        ref = parity.oracle_hmc(model, q0, 0.4, sigma ** 2, draws, T, L)
        got = hostsim.run(0, mu, model.inv_var, 0.0, sigma ** 2, q0, 0.4, draws, T, hmc_L=L, n_store=T, reg_front=True)
        for k in ("q", "p", "g", "U", "acceptance_probability"):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-11, atol=1e-13, err_msg=k)


@pytest.mark.parametrize("model_name, eps", [("iid", 0.4), ("iid", 1.7), ("funnel", 0.3)])
def test_exact_doubling_engine_matches_oracle(model_name, eps):
    """The non-reference option `exact_doubling` (sub-trees of 2**k leapfrogs): the engine code and the oracle's
    variant agree decision for decision, and the trajectories really are balanced (2**depth - 1 leapfrogs when
    nothing stopped a sub-tree early)."""
    rng = np.random.default_rng(31)
    C, T = 16, 3
    if model_name == "iid":
        d = 5
        mu, sigma = rng.standard_normal(d), np.exp(0.5 * rng.standard_normal(d))
        model, args = models.IIDGaussian(mu, sigma), (0, mu, 1.0 / sigma ** 2, 0.0)
        q0 = mu + sigma * rng.standard_normal((C, d))
    else:
        d = 10
        model, args = models.NealFunnel(d), (2, None, None, 0.0)
        q0 = rng.standard_normal((C, d))
    imm = np.ones(d)
    draws = parity.random_draws(rng, C, T, d)
    ref = parity.oracle_nuts(model, q0, eps, imm, draws, T, exact_doubling=True)
    got = hostsim.run(*args, imm, q0, eps, draws, T, n_store=T, exact_doubling=True)
    parity.assert_nuts_parity(got, ref, rtol=1e-10, what=f"exact doubling {model_name}")
    ref_q = parity.oracle_nuts(model, q0, eps, imm, draws, T)          # the reference's 2**k + 1 behaviour differs
    assert not np.array_equal(ref_q["n_leapfrog"], ref["n_leapfrog"])
    clean = ~ref["is_diverging"].astype(bool)
    assert np.all(ref["n_leapfrog"][clean] <= 2 ** ref["num_doublings"][clean] - 1)


def test_exact_doubling_removes_the_reference_bias():
    """Oracle-level statement of DESIGN.md 2.1 on the reference's own 1-d test target, logprob = -2 (x - 1)^2
    (tests/test_step_size.py:15-16; true variance 1/4) at eps = 0.5: the reference's 2**k + 1 sub-trees resonate with
    the oscillator and give variance ~0.002, balanced sub-trees give the posterior."""
    from oracle import kernels, streams

    class Quad:
        def potential_and_grad(self, q):
            r = q - 1.0
            return float(2.0 * np.sum(r * r)), 4.0 * r

    def long_run_variance(exact):
        xs = []
        for seed in range(24):
            srng = streams.StreamDraws(seed, "nuts")
            k = kernels.nuts_new_kernel(srng, Quad(), exact_doubling=exact)
            st = kernels.new_state(np.zeros(1), Quad())
            for t in range(70):
                info, _ = k(st, 0.5, np.ones(1))
                st = info.state._replace(momentum=None)
                if t >= 20:
                    xs.append(info.state.position[0])
        return np.var(xs)

    assert long_run_variance(False) < 0.02
    assert 0.19 < long_run_variance(True) < 0.31
