"""Pins the oracle to everything the reference publishes for the trajectory path.

* README quick-start draw (reference README.md:22-55) -- bit-exact.
* tests/test_trajectory.py:144-208 (multiplicative expansion flags, seed 59).
* SURVEY.md section 3.7 trace of the README transition.
"""
import numpy as np
import pytest

from oracle import hamiltonian, kernels, models, streams, tree


def test_readme_quickstart_bit_exact():
    srng = streams.StreamDraws(0, "nuts")
    model = models.Normal(0.0, 1.0)
    kernel = kernels.nuts_new_kernel(srng, model)
    state = kernels.new_state(np.float64(0.0), model)
    info, extras = kernel(state, 1e-2, np.float64(1.0))
    assert float(info.state.position) == 1.1034719409361107   # reference README.md:54
    assert info.num_doublings == 8
    assert extras["n_leapfrog"] == 136
    assert info.is_turning is False and info.is_diverging is False
    lengths = [t.subtree_length for t in extras["trace"]]
    assert lengths == [2, 3, 5, 9, 17, 33, 65, 2]             # Q1: 2**k + 1 leapfrogs
    assert [int(t.go_right) for t in extras["trace"]] == [1, 0, 1, 0, 1, 1, 1, 1]
    assert extras["trace"][-1].subtree_terminated             # Q2: stale checkpoint index


def test_stream_topology_first_draws():
    # first momentum normal of seed 0 / seed 59
    assert streams.StreamDraws(0).normal(()) == 1.4436909546981256
    s59 = streams.StreamDraws(59)
    assert s59.normal(()) == -0.29245959035980823
    assert s59.direction(0) is False


def test_bernoulli_rule_matches_numpy_binomial():
    """``Generator.binomial(1, p)`` consumes one double u and applies
    ``bernoulli_from_uniform`` (the rule the CUDA validation mode uses)."""
    rng = np.random.default_rng(1234)
    ps = np.concatenate([rng.random(500), [0.5, 1.0, 1e-300, 0.5 + 1e-16, 0.999999]])
    for p in ps:
        bg = np.random.PCG64(int(rng.integers(1 << 30)))
        clone = np.random.PCG64()
        clone.state = bg.state
        got = np.random.Generator(bg).binomial(1, p)
        u = np.random.Generator(clone).random()
        assert bool(got) == streams.bernoulli_from_uniform(u, p), (p, u)
    # p == 0 consumes nothing
    bg = np.random.PCG64(7)
    before = bg.state["state"]["state"]
    assert np.random.Generator(bg).binomial(1, 0.0) == 0
    assert bg.state["state"]["state"] == before


def _expansion_case(step_size):
    """reference tests/test_trajectory.py:144-208, streams M,D,U,B = children 0..3 of seed 59."""
    srng = streams.StreamDraws(59, "nuts")
    potential = lambda x: (0.5 * np.sum(np.square(x)), x)
    imm = np.float64(1.0)
    position = np.float64(1.0)
    momentum_generator, kinetic_energy_fn, uturn_check_fn = hamiltonian.gaussian_metric(imm)
    integrator = hamiltonian.velocity_verlet(potential, kinetic_energy_fn)
    new_criterion_state, update_criterion_state, is_criterion_met = tree.iterative_uturn(uturn_check_fn)
    trajectory_integrator = tree.dynamic_integration(
        srng, integrator, kinetic_energy_fn, update_criterion_state, is_criterion_met, 1000)
    expand = tree.multiplicative_expansion(srng, trajectory_integrator, uturn_check_fn, 10)
    state = hamiltonian.new_integrator_state(potential, position, momentum_generator(srng))
    energy = state.potential_energy + kinetic_energy_fn(state.momentum)
    proposal = tree.ProposalState(state, energy, 0.0, -np.inf)
    termination_state = new_criterion_state(state.position, 10)
    return expand(proposal, state, state, state.momentum, termination_state, energy, step_size)


@pytest.mark.parametrize(
    "step_size, should_diverge, should_turn, expected_doublings",
    [(100000.0, True, False, 1), (0.0000001, False, False, 10), (1.0, False, True, 1)],
)
def test_multiplicative_expansion(step_size, should_diverge, should_turn, expected_doublings):
    info, extras = _expansion_case(step_size)
    assert info.is_diverging == should_diverge
    assert info.is_turning == should_turn
    assert info.num_doublings == expected_doublings


def test_multiplicative_expansion_regression_values():
    # SURVEY.md 3.7 (emulation written independently during the survey)
    info, extras = _expansion_case(1e-7)
    assert [t.subtree_length for t in extras["trace"]] == [2, 3, 5, 9, 17, 33, 65, 129, 257, 513]
    assert extras["n_leapfrog"] == 1033
    assert [int(t.go_right) for t in extras["trace"]] == [0, 1, 1, 1, 0, 1, 0, 0, 0, 0]
    assert float(info.state.position) == pytest.approx(1.0000287147101077, rel=1e-14)
    info, extras = _expansion_case(1.0)
    assert extras["trace"][0].subtree_length == 2 and not extras["trace"][0].go_right
    assert float(info.state.position) == pytest.approx(-0.20754040964019183, rel=1e-14)
    info, extras = _expansion_case(1e5)
    assert extras["trace"][0].subtree_length == 1


@pytest.mark.parametrize("case", [(0.0000001, False, False), (1000, True, False), (1e100, True, False)])
def test_dynamic_integration(case):
    """reference tests/test_trajectory.py:77-141 (flags only; N(0,1) potential)."""
    step_size, should_diverge, should_turn = case
    srng = streams.StreamDraws(59, "nuts")
    model = models.Normal(0.0, 1.0)
    momentum_generator, kinetic_energy_fn, uturn_check_fn = hamiltonian.gaussian_metric(np.ones(1))
    integrator = hamiltonian.velocity_verlet(model.potential_and_grad, kinetic_energy_fn)
    new_criterion_state, update_criterion_state, is_criterion_met = tree.iterative_uturn(uturn_check_fn)
    trajectory_integrator = tree.dynamic_integration(
        srng, integrator, kinetic_energy_fn, update_criterion_state, is_criterion_met, 1000)
    initial_state = hamiltonian.new_integrator_state(
        model.potential_and_grad, np.ones(1), momentum_generator(srng))
    initial_energy = initial_state[2] + kinetic_energy_fn(initial_state[1])
    termination_state = new_criterion_state(initial_state[0], 10)
    with np.errstate(all="ignore"):
        sub, _ = trajectory_integrator(initial_state, 1, termination_state, 10, step_size, initial_energy)
    assert sub.is_diverging is should_diverge
    assert sub.has_terminated is should_turn
