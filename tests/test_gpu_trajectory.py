"""GPU parity of the stand-alone trajectory builders and proposal primitives (b2h_nuts_subtree,
b2h_nuts_expand, b2h_proposal_update, b2h_progressive_sampling, b2h_select_rows) against the oracle,
plus the reference's own fixtures (reference tests/test_trajectory.py:72-222, tests/test_proposals.py)."""
import numpy as np
import pytest

import parity
from oracle import hamiltonian as o_ham
from oracle import models as o_models
from oracle import tree as o_tree

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ab(cuda_device):
    import aehmc_b200
    return aehmc_b200


def _np(t):
    return t.detach().cpu().numpy()


def _oracle_parts(om, imm, srng, thr=1000.0):
    mg, ke, ut = o_ham.gaussian_metric(imm)
    integ = o_ham.velocity_verlet(om.potential_and_grad, ke)
    new_ts, upd, crit = o_tree.iterative_uturn(ut)
    ti = o_tree.dynamic_integration(srng, integ, ke, upd, crit, thr)
    return mg, ke, ut, new_ts, ti


def _gpu_parts(ab, gm, imm, srng, thr=1000.0, expansion=None, group=0):
    mg, ke, ut = ab.metrics.gaussian_metric(imm, dtype=gm.dtype)
    integ = ab.integrators.velocity_verlet(gm, ke)
    new_ts, upd, crit = ab.termination.iterative_uturn(ut)
    ti = ab.trajectory.dynamic_integration(srng, integ, ke, upd, crit, thr, expansion=expansion, group=group)
    return mg, ke, ut, new_ts, ti


def _setup(rng, C, d):
    mu, sigma = 0.3 * rng.standard_normal(d), np.exp(0.3 * rng.standard_normal(d))
    imm = np.exp(0.2 * rng.standard_normal(d))
    q0 = mu + sigma * rng.standard_normal((C, d))
    p0 = rng.standard_normal((C, d)) / np.sqrt(imm)
    return mu, sigma, imm, q0, p0


@pytest.mark.parametrize("group,d", [(1, 5), (8, 5), (8, 40), (32, 70)])
@pytest.mark.parametrize("max_steps,expansion", [(8, 3), (10, 3), (1, 0)])
def test_subtree_matches_oracle(ab, group, d, max_steps, expansion):
    """one sub-tree per chain from a caller-supplied state, both directions, vs oracle tree.dynamic_integration."""
    rng = np.random.default_rng(100 + d + max_steps)
    C, maxd = 24, 6
    mu, sigma, imm, q0, p0 = _setup(rng, C, d)
    eps = np.where(np.arange(C) % 5 == 4, 30.0, 0.35 + 0.1 * rng.random(C))      # every fifth chain diverges
    eps[::7] = 1.1                                                               # some chains U-turn inside the sub-tree
    dirs = np.where(rng.random(C) < 0.5, 1, -1).astype(np.int8)
    draws = parity.random_draws(rng, C, 1, d, maxd)
    om = o_models.IIDGaussian(mu, sigma)
    gm = ab.models.IIDGaussian(mu, sigma)
    # termination state with non-trivial content (as left by earlier expansions)
    mck0 = rng.standard_normal((C, maxd, d)); sck0 = rng.standard_normal((C, maxd, d))

    ref = []
    for c in range(C):
        srng = parity.chain_draws(draws, c)
        srng.begin_transition()
        mg, ke, ut, new_ts, ti = _oracle_parts(om, imm, srng)
        U, g = om.potential_and_grad(q0[c])
        st = o_ham.IntegratorState(q0[c], p0[c], U, g)
        E0 = U + ke(p0[c])
        ts = o_tree.TerminationState(mck0[c].copy(), sck0[c].copy(), 0, 0)
        with np.errstate(all="ignore"):
            sub, _ = ti(st, float(dirs[c]), ts, max_steps, float(eps[c]), E0, expansion=expansion)
        ref.append((sub, E0))

    srng = ab.InjectedDraws(None, draws["u_dir"], draws["u_biased"], draws["u_uniform"], None)
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, imm, srng, expansion=expansion, group=group)
    U, g = gm.potential_and_grad(q0)
    st = ab.integrators.IntegratorState(torch.as_tensor(q0).cuda(), torch.as_tensor(p0).cuda(), U, g)
    E0 = U + ke(st.momentum)
    ts = ab.termination.TerminationState(torch.as_tensor(mck0).cuda(), torch.as_tensor(sck0).cuda(),
                                         torch.zeros(C, dtype=torch.int64).cuda(), torch.zeros(C, dtype=torch.int64).cuda())
    (prop, last, msum, ts2, length, div, term), _ = ti(st, dirs, ts, max_steps, eps, E0)

    np.testing.assert_array_equal(_np(length), [r[0].trajectory_length for r in ref])
    np.testing.assert_array_equal(_np(div), [bool(r[0].is_diverging) for r in ref])
    np.testing.assert_array_equal(_np(term), [bool(r[0].has_terminated) for r in ref])
    assert _np(div).any() and _np(term).any() and (~_np(div) & ~_np(term)).any() or max_steps == 1
    np.testing.assert_array_equal(_np(ts2.min_index), [r[0].termination_state.min_index for r in ref])
    np.testing.assert_array_equal(_np(ts2.max_index), [r[0].termination_state.max_index for r in ref])
    ok = ~_np(div)                      # positions of divergent chains overflow; flags and lengths are the contract
    tol = dict(rtol=1e-10, atol=1e-11)
    for name, got, want in [
        ("proposal.q", prop.state.position, [r[0].proposal.state.position for r in ref]),
        ("proposal.p", prop.state.momentum, [r[0].proposal.state.momentum for r in ref]),
        ("proposal.g", prop.state.potential_energy_grad, [r[0].proposal.state.potential_energy_grad for r in ref]),
        ("proposal.U", prop.state.potential_energy, [r[0].proposal.state.potential_energy for r in ref]),
        ("proposal.energy", prop.energy, [r[0].proposal.energy for r in ref]),
        ("proposal.weight", prop.weight, [r[0].proposal.weight for r in ref]),
        ("proposal.slpa", prop.sum_log_p_accept, [r[0].proposal.sum_log_p_accept for r in ref]),
        ("last.q", last.position, [r[0].state.position for r in ref]),
        ("last.p", last.momentum, [r[0].state.momentum for r in ref]),
        ("last.U", last.potential_energy, [r[0].state.potential_energy for r in ref]),
        ("last.g", last.potential_energy_grad, [r[0].state.potential_energy_grad for r in ref]),
        ("momentum_sum", msum, [r[0].momentum_sum for r in ref]),
        ("mck", ts2.momentum_checkpoints, [r[0].termination_state.momentum_checkpoints for r in ref]),
        ("sck", ts2.momentum_sum_checkpoints, [r[0].termination_state.momentum_sum_checkpoints for r in ref]),
    ]:
        np.testing.assert_allclose(_np(got)[ok], np.asarray(want, dtype=np.float64)[ok], err_msg=name, **tol)


@pytest.mark.parametrize("group,d", [(1, 4), (8, 4), (8, 33), (32, 90)])
def test_expand_matches_oracle(ab, group, d):
    """the whole doubling loop from a caller-supplied tree vs oracle tree.multiplicative_expansion."""
    rng = np.random.default_rng(200 + d)
    C, maxd = 32, 7
    mu, sigma, imm, q0, p0 = _setup(rng, C, d)
    eps = np.where(np.arange(C) % 8 == 7, 25.0, 0.2 + 0.3 * rng.random(C))
    eps[::5] = 0.02                                                    # runs to the expansion cap
    draws = parity.random_draws(rng, C, 1, d, maxd)
    om = o_models.IIDGaussian(mu, sigma)
    gm = ab.models.IIDGaussian(mu, sigma)

    ref = []
    for c in range(C):
        srng = parity.chain_draws(draws, c)
        srng.begin_transition()
        mg, ke, ut, new_ts, ti = _oracle_parts(om, imm, srng)
        expand = o_tree.multiplicative_expansion(srng, ti, ut, maxd)
        U, g = om.potential_and_grad(q0[c])
        st = o_ham.IntegratorState(q0[c], p0[c], U, g)
        E0 = U + ke(p0[c])
        with np.errstate(all="ignore"):
            info, ex = expand(o_tree.ProposalState(st, E0, 0.0, -np.inf), st, st, st.momentum,
                              new_ts(st.position, maxd), E0, float(eps[c]))
        ref.append((info, ex))

    srng = ab.InjectedDraws(None, draws["u_dir"], draws["u_biased"], draws["u_uniform"], None)
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, imm, srng, group=group)
    expand = ab.trajectory.multiplicative_expansion(srng, ti, ut, maxd)
    U, g = gm.potential_and_grad(q0)
    st = ab.integrators.IntegratorState(torch.as_tensor(q0).cuda(), torch.as_tensor(p0).cuda(), U, g)
    E0 = U + ke(st.momentum)
    prop = ab.proposals.ProposalState(st, E0, torch.zeros(C, dtype=torch.float64).cuda(),
                                      torch.full((C,), -np.inf, dtype=torch.float64).cuda())
    res, _ = expand(prop, st, st, st.momentum, new_ts(st.position, maxd), E0, eps)

    dg = res.diagnostics
    np.testing.assert_array_equal(_np(dg.num_doublings[-1]), [r[0].num_doublings for r in ref])
    np.testing.assert_array_equal(_np(dg.is_turning[-1]), [bool(r[0].is_turning) for r in ref])
    np.testing.assert_array_equal(_np(dg.is_diverging[-1]), [bool(r[0].is_diverging) for r in ref])
    nd = _np(dg.num_doublings[-1])
    assert nd.max() == maxd and nd.min() == 1 and _np(dg.is_turning[-1]).any() and _np(dg.is_diverging[-1]).any()
    ok = ~_np(dg.is_diverging[-1])
    tol = dict(rtol=1e-10, atol=1e-11)
    for name, got, want in [
        ("acceptance", dg.acceptance_probability[-1], [r[0].acceptance_probability for r in ref]),
        ("proposal.q", res.proposals.state.position[-1], [r[1]["proposal"].state.position for r in ref]),
        ("proposal.p", res.proposals.state.momentum[-1], [r[1]["proposal"].state.momentum for r in ref]),
        ("proposal.U", res.proposals.state.potential_energy[-1], [r[1]["proposal"].state.potential_energy for r in ref]),
        ("proposal.energy", res.proposals.energy[-1], [r[1]["proposal"].energy for r in ref]),
        ("proposal.weight", res.proposals.weight[-1], [r[1]["proposal"].weight for r in ref]),
        ("proposal.slpa", res.proposals.sum_log_p_accept[-1], [r[1]["proposal"].sum_log_p_accept for r in ref]),
        ("left.q", res.left_states.position[-1], [r[1]["left_state"].position for r in ref]),
        ("left.p", res.left_states.momentum[-1], [r[1]["left_state"].momentum for r in ref]),
        ("left.U", res.left_states.potential_energy[-1], [r[1]["left_state"].potential_energy for r in ref]),
        ("right.q", res.right_states.position[-1], [r[1]["right_state"].position for r in ref]),
        ("right.g", res.right_states.potential_energy_grad[-1], [r[1]["right_state"].potential_energy_grad for r in ref]),
        ("right.U", res.right_states.potential_energy[-1], [r[1]["right_state"].potential_energy for r in ref]),
        ("momentum_sum", res.momentum_sums[-1], [r[1]["momentum_sum"] for r in ref]),
        ("mck", res.termination_states.momentum_checkpoints[-1], [r[1]["termination_state"].momentum_checkpoints for r in ref]),
    ]:
        np.testing.assert_allclose(_np(got)[ok], np.asarray(want, dtype=np.float64)[ok], err_msg=name, **tol)
    np.testing.assert_array_equal(_np(res.termination_states.min_index[-1]), [r[1]["termination_state"].min_index for r in ref])
    np.testing.assert_array_equal(_np(res.termination_states.max_index[-1]), [r[1]["termination_state"].max_index for r in ref])


def test_expand_equals_nuts_kernel(ab):
    """expand() entered with the state nuts.new_kernel builds gives the kernel's own transition."""
    rng = np.random.default_rng(5)
    C, d, maxd = 64, 10, 8
    q0 = rng.standard_normal((C, d))
    gm = ab.models.NealFunnel(d)
    draws = parity.random_draws(rng, C, 1, d, maxd)
    srng = ab.InjectedDraws(**draws)
    kernel = ab.nuts.new_kernel(srng, gm, max_num_expansions=maxd)
    state = ab.nuts.new_state(q0, gm)
    imm = np.ones(d)
    info, _ = kernel(state, 0.2, imm)

    srng2 = ab.InjectedDraws(**draws)
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, imm, srng2)
    expand = ab.trajectory.multiplicative_expansion(srng2, ti, ut, maxd)
    p0 = mg(srng2, C, d, transition=0)
    st = ab.integrators.IntegratorState(state.position, p0, state.potential_energy, state.potential_energy_grad)
    E0 = st.potential_energy + ke(p0)
    prop = ab.proposals.ProposalState(st, E0, torch.zeros(C, dtype=torch.float64).cuda(),
                                      torch.full((C,), -np.inf, dtype=torch.float64).cuda())
    res, _ = expand(prop, st, st, p0, new_ts(st.position, maxd), E0, 0.2)
    np.testing.assert_array_equal(_np(res.diagnostics.num_doublings[-1]), _np(info.num_doublings))
    np.testing.assert_array_equal(_np(res.diagnostics.is_turning[-1]), _np(info.is_turning))
    np.testing.assert_allclose(_np(res.diagnostics.state.position[-1]), _np(info.state.position), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(_np(res.diagnostics.acceptance_probability[-1]), _np(info.acceptance_probability), rtol=1e-12)


# ---- the reference's own fixtures ------------------------------------------------------------
@pytest.mark.parametrize("step_size,should_diverge,should_turn",
                         [(0.0000001, False, False), (1000, True, False), (1e100, True, False)])
def test_dynamic_integration_reference_cases(ab, step_size, should_diverge, should_turn):
    """reference tests/test_trajectory.py:72-141 (standard normal target, unit metric, 10 steps)."""
    srng = ab.RandomStream(seed=59)
    gm = ab.models.IIDGaussian(np.zeros(1), np.ones(1))
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, np.ones(1), srng)
    position = torch.ones((3, 1), dtype=torch.float64).cuda()
    state = ab.integrators.new_integrator_state(gm, position, mg(srng, 3, 1))
    E0 = state.potential_energy + ke(state.momentum)
    ts = new_ts(state.position, 10)
    out, _ = ti(state, 1, ts, 10, step_size, E0)
    assert _np(out[-2]).tolist() == [should_diverge] * 3
    assert _np(out[-1]).tolist() == [should_turn] * 3


@pytest.mark.parametrize("step_size,should_diverge,should_turn,expected_doublings",
                         [(100000.0, True, False, 1), (0.0000001, False, False, 10), (1.0, False, True, 1)])
def test_multiplicative_expansion_reference_cases(ab, step_size, should_diverge, should_turn, expected_doublings):
    """reference tests/test_trajectory.py:144-222 (U = q^2/2, unit metric, position 1)."""
    srng = ab.RandomStream(seed=59)
    gm = ab.models.IIDGaussian(np.zeros(1), np.ones(1), const=-0.5 * np.log(2 * np.pi))
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, 1.0, srng)
    expand = ab.trajectory.multiplicative_expansion(srng, ti, ut, 10)
    C = 5
    position = torch.ones((C, 1), dtype=torch.float64).cuda()
    state = ab.integrators.new_integrator_state(gm, position, mg(srng, C, 1))
    energy = state.potential_energy + ke(state.momentum)
    prop = ab.proposals.ProposalState(state, energy, torch.zeros(C, dtype=torch.float64).cuda(),
                                      torch.full((C,), -np.inf, dtype=torch.float64).cuda())
    res, _ = expand(prop, state, state, state.momentum, new_ts(state.position, 10), energy, step_size)
    # the reference asserts these for its one seeded momentum; over several momenta the divergence flag and
    # the doubling count of the diverging / tiny-step cases are draw-independent, the U-turn of the eps = 1
    # case (a 60-degree rotation per leapfrog) lands within the first two doublings
    assert _np(res.diagnostics.is_diverging[-1]).tolist() == [should_diverge] * C
    nd = _np(res.diagnostics.num_doublings[-1])
    if should_turn:
        assert _np(res.diagnostics.is_turning[-1]).all() and nd.min() == expected_doublings and nd.max() <= 2
    else:
        assert nd.tolist() == [expected_doublings] * C
        if not should_diverge:
            assert not _np(res.diagnostics.is_turning[-1]).any()


def test_proposal_primitives(ab):
    """proposals.py:41-52, 96-100, 130-174 vs the oracle, incl. NaN / inf energies."""
    rng = np.random.default_rng(9)
    C, d = 200, 3
    imm = np.ones(d)
    mg, ke, ut = ab.metrics.gaussian_metric(imm)
    omg, oke, out_ = o_ham.gaussian_metric(imm)
    q, p = rng.standard_normal((C, d)), rng.standard_normal((C, d))
    U = rng.standard_normal(C) * 3
    U[::17] = np.nan; U[5::23] = np.inf; U[7::29] = 2000.0
    g = rng.standard_normal((C, d))
    E0 = rng.standard_normal(C)
    cu = lambda a: torch.as_tensor(a).cuda()
    gen = ab.proposals.proposal_generator(ke, 1000.0)
    st = ab.integrators.IntegratorState(cu(q), cu(p), cu(U), cu(g))
    prop, div = gen(cu(E0), st)
    ogen = o_tree.proposal_generator(oke, 1000.0)
    with np.errstate(all="ignore"):
        ref = [ogen(E0[c], o_ham.IntegratorState(q[c], p[c], U[c], g[c])) for c in range(C)]
    np.testing.assert_array_equal(_np(div), [bool(r[1]) for r in ref])
    np.testing.assert_allclose(_np(prop.energy), [r[0].energy for r in ref], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(_np(prop.weight), [r[0].weight for r in ref], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(_np(prop.sum_log_p_accept), [r[0].sum_log_p_accept for r in ref], rtol=1e-12, atol=1e-14)

    # progressive sampling between two proposals with injected uniforms
    w1, w2 = rng.standard_normal(C) * 4, rng.standard_normal(C) * 4
    w2[::11] = -np.inf; w1[3::13] = -np.inf
    s1, s2 = -rng.random(C) * 5, -rng.random(C) * 5
    s1[::9] = -np.inf
    u = rng.random((C, 2))
    mk = lambda w, s, shift: ab.proposals.ProposalState(
        ab.integrators.IntegratorState(cu(q + shift), cu(p + shift), cu(E0 + shift), cu(g + shift)), cu(E0 - shift), cu(w), cu(s))
    a, b = mk(w1, s1, 0.0), mk(w2, s2, 1.0)
    srng = ab.InjectedDraws(None, u_accept=u)
    for k, (fn, ofn) in enumerate([(ab.proposals.progressive_uniform_sampling, "uniform"),
                                   (ab.proposals.progressive_biased_sampling, "biased")]):
        got = fn(srng, a, b)
        for c in range(C):
            if ofn == "uniform":
                pa = o_tree.expit(w2[c] - w1[c]); pa = 0.0 if np.isnan(pa) else pa
            else:
                with np.errstate(all="ignore"):
                    pa = min(max(np.exp(w2[c] - w1[c]), 0.0), 1.0)
            from oracle.streams import bernoulli_from_uniform
            take = bernoulli_from_uniform(u[c, k], pa) if not np.isnan(pa) else False
            want_q = q[c] + (1.0 if take else 0.0)
            np.testing.assert_array_equal(_np(got.state.position[c]), want_q, err_msg=f"{ofn} chain {c}")
            assert _np(got.energy[c]) == E0[c] - (1.0 if take else 0.0)
            np.testing.assert_allclose(_np(got.weight[c]), o_tree.logaddexp(w1[c], w2[c]), rtol=1e-14)
            np.testing.assert_allclose(_np(got.sum_log_p_accept[c]), o_tree.logaddexp(s1[c], s2[c]), rtol=1e-14)

    mask = rng.random(C) < 0.5
    sel = ab.trajectory.where_proposal(cu(mask), a, b)
    np.testing.assert_array_equal(_np(sel.state.position), np.where(mask[:, None], q, q + 1.0))
    np.testing.assert_array_equal(_np(sel.weight), np.where(mask, w1, w2))
    upd = ab.proposals.maybe_update_proposal(cu(mask), a, b)
    np.testing.assert_array_equal(_np(upd.state.momentum), np.where(mask[:, None], p + 1.0, p))


@pytest.mark.parametrize("kind", ["corr", "logistic"])
def test_standalone_builders_split_models(ab, kind):
    """dynamic_integration / multiplicative_expansion as stand-alone closures for models that run in the per-tick
    (split) engine: correlated Gaussian and logistic regression with a diagonal metric."""
    rng = np.random.default_rng(77)
    C, maxd, d = 20, 6, 6
    if kind == "corr":
        A = rng.standard_normal((d, d))
        cov = A @ A.T / d + 0.3 * np.eye(d)
        mu = 0.2 * rng.standard_normal(d)
        om = o_models.CorrelatedGaussian(mu, np.linalg.inv(cov))
        gm = ab.models.CorrelatedGaussian(mu, np.linalg.inv(cov))
    else:
        X = rng.standard_normal((80, d)); y = (rng.random(80) < 0.5).astype(np.float64)
        om = o_models.LogisticRegression(X, y, 1.0)
        gm = ab.models.LogisticRegression(X, y, 1.0)
    imm = np.exp(0.2 * rng.standard_normal(d))
    q0 = 0.5 * rng.standard_normal((C, d))
    p0 = rng.standard_normal((C, d)) / np.sqrt(imm)
    eps = 0.15 + 0.2 * rng.random(C)
    eps[::6] = 20.0                                                    # some chains diverge
    draws = parity.random_draws(rng, C, 1, d, maxd)

    # --- the whole doubling loop
    ref = []
    for c in range(C):
        srng = parity.chain_draws(draws, c)
        srng.begin_transition()
        mg, ke, ut, new_ts, ti = _oracle_parts(om, imm, srng)
        expand = o_tree.multiplicative_expansion(srng, ti, ut, maxd)
        U, g = om.potential_and_grad(q0[c])
        st = o_ham.IntegratorState(q0[c], p0[c], U, g)
        E0 = U + ke(p0[c])
        with np.errstate(all="ignore"):
            ref.append(expand(o_tree.ProposalState(st, E0, 0.0, -np.inf), st, st, st.momentum, new_ts(st.position, maxd), E0,
                              float(eps[c])))
    srng = ab.InjectedDraws(None, draws["u_dir"], draws["u_biased"], draws["u_uniform"], None)
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, imm, srng)
    expand = ab.trajectory.multiplicative_expansion(srng, ti, ut, maxd)
    U, g = gm.potential_and_grad(q0)
    st = ab.integrators.IntegratorState(torch.as_tensor(q0).cuda(), torch.as_tensor(p0).cuda(), U, g)
    E0 = U + ke(st.momentum)
    prop = ab.proposals.ProposalState(st, E0, torch.zeros(C, dtype=torch.float64).cuda(),
                                      torch.full((C,), -np.inf, dtype=torch.float64).cuda())
    res, _ = expand(prop, st, st, st.momentum, new_ts(st.position, maxd), E0, eps)
    dg = res.diagnostics
    np.testing.assert_array_equal(_np(dg.num_doublings[-1]), [r[0].num_doublings for r in ref])
    np.testing.assert_array_equal(_np(dg.is_turning[-1]), [bool(r[0].is_turning) for r in ref])
    np.testing.assert_array_equal(_np(dg.is_diverging[-1]), [bool(r[0].is_diverging) for r in ref])
    ok = ~_np(dg.is_diverging[-1])
    assert ok.any() and (~ok).any()
    for name, got, want in [
        ("proposal.q", res.proposals.state.position[-1], [r[1]["proposal"].state.position for r in ref]),
        ("left.q", res.left_states.position[-1], [r[1]["left_state"].position for r in ref]),
        ("right.g", res.right_states.potential_energy_grad[-1], [r[1]["right_state"].potential_energy_grad for r in ref]),
        ("momentum_sum", res.momentum_sums[-1], [r[1]["momentum_sum"] for r in ref]),
        ("acceptance", dg.acceptance_probability[-1], [r[0].acceptance_probability for r in ref]),
    ]:
        np.testing.assert_allclose(_np(got)[ok], np.asarray(want, dtype=np.float64)[ok], err_msg=name, rtol=1e-9, atol=1e-11)

    # --- one sub-tree
    dirs = np.where(rng.random(C) < 0.5, 1, -1).astype(np.int8)
    refs = []
    for c in range(C):
        srng = parity.chain_draws(draws, c)
        srng.begin_transition()
        mg_o, ke_o, ut_o, new_ts_o, ti_o = _oracle_parts(om, imm, srng)
        U, g = om.potential_and_grad(q0[c])
        st_o = o_ham.IntegratorState(q0[c], p0[c], U, g)
        with np.errstate(all="ignore"):
            sub, _ = ti_o(st_o, float(dirs[c]), new_ts_o(st_o.position, maxd), 4, float(eps[c]), U + ke_o(p0[c]), expansion=2)
        refs.append(sub)
    srng = ab.InjectedDraws(None, draws["u_dir"], draws["u_biased"], draws["u_uniform"], None)
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, imm, srng, expansion=2)
    (prop2, last, msum, ts2, length, div, term), _ = ti(st, dirs, new_ts(st.position, maxd), 4, eps, E0)
    np.testing.assert_array_equal(_np(length), [r.trajectory_length for r in refs])
    np.testing.assert_array_equal(_np(div), [bool(r.is_diverging) for r in refs])
    ok = ~_np(div)
    np.testing.assert_allclose(_np(last.position)[ok], np.asarray([r.state.position for r in refs])[ok], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(_np(prop2.state.position)[ok], np.asarray([r.proposal.state.position for r in refs])[ok],
                               rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("group", [1, 8])
def test_expand_with_per_chain_diagonal_metric(ab, group):
    """every chain with its own diagonal inverse mass matrix (what per-chain window adaptation produces)."""
    rng = np.random.default_rng(91)
    C, maxd, d = 16, 6, 5
    mu, sigma = 0.3 * rng.standard_normal(d), np.exp(0.3 * rng.standard_normal(d))
    imm = np.exp(0.4 * rng.standard_normal((C, d)))
    q0 = mu + sigma * rng.standard_normal((C, d))
    p0 = rng.standard_normal((C, d)) / np.sqrt(imm)
    eps = 0.2 + 0.3 * rng.random(C)
    draws = parity.random_draws(rng, C, 1, d, maxd)
    om = o_models.IIDGaussian(mu, sigma)
    gm = ab.models.IIDGaussian(mu, sigma)
    ref = []
    for c in range(C):
        srng = parity.chain_draws(draws, c)
        srng.begin_transition()
        mg, ke, ut, new_ts, ti = _oracle_parts(om, imm[c], srng)
        expand = o_tree.multiplicative_expansion(srng, ti, ut, maxd)
        U, g = om.potential_and_grad(q0[c])
        st = o_ham.IntegratorState(q0[c], p0[c], U, g)
        E0 = U + ke(p0[c])
        ref.append(expand(o_tree.ProposalState(st, E0, 0.0, -np.inf), st, st, st.momentum, new_ts(st.position, maxd), E0,
                          float(eps[c])))
    srng = ab.InjectedDraws(None, draws["u_dir"], draws["u_biased"], draws["u_uniform"], None)
    mg, ke, ut, new_ts, ti = _gpu_parts(ab, gm, ab.metrics.per_chain(torch.as_tensor(imm).cuda()), srng, group=group)
    expand = ab.trajectory.multiplicative_expansion(srng, ti, ut, maxd)
    U, g = gm.potential_and_grad(q0)
    st = ab.integrators.IntegratorState(torch.as_tensor(q0).cuda(), torch.as_tensor(p0).cuda(), U, g)
    E0 = U + ke(st.momentum)
    prop = ab.proposals.ProposalState(st, E0, torch.zeros(C, dtype=torch.float64).cuda(),
                                      torch.full((C,), -np.inf, dtype=torch.float64).cuda())
    res, _ = expand(prop, st, st, st.momentum, new_ts(st.position, maxd), E0, eps)
    np.testing.assert_array_equal(_np(res.diagnostics.num_doublings[-1]), [r[0].num_doublings for r in ref])
    np.testing.assert_array_equal(_np(res.diagnostics.is_turning[-1]), [bool(r[0].is_turning) for r in ref])
    np.testing.assert_allclose(_np(res.proposals.state.position[-1]), [r[1]["proposal"].state.position for r in ref],
                               rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(_np(res.momentum_sums[-1]), [r[1]["momentum_sum"] for r in ref], rtol=1e-10, atol=1e-11)
