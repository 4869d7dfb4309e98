"""The reference's exact/analytic unit fixtures, replayed on the oracle.

Sources: reference tests/test_termination.py, test_metrics.py, test_integrators.py,
test_trajectory.py:33-74, test_algorithms.py, test_adaptation.py.
"""
import numpy as np
import pytest

from oracle import adaptation, hamiltonian, tree
from oracle.hamiltonian import IntegratorState


# ---- termination (reference tests/test_termination.py) ----------------------
@pytest.mark.parametrize(
    "checkpoint_idxs, momentum, momentum_sum, inverse_mass_matrix, expected_turning",
    [
        ((3, 3), 1.0, 3.0, 1.0, True),
        ((3, 2), 1.0, 3.0, 1.0, False),
        ((0, 0), 1.0, 3.0, 1.0, False),
        ((0, 1), 1.0, 3.0, 1.0, True),
        ((1, 3), 1.0, 3.0, 1.0, True),
        ((1, 3), np.array([1.0]), np.array([3.0]), np.ones(1), True),
    ],
)
def test_iterative_turning_termination(checkpoint_idxs, momentum, momentum_sum, inverse_mass_matrix,
                                       expected_turning):
    _, _, is_turning = hamiltonian.gaussian_metric(inverse_mass_matrix)
    _, _, is_iterative_turning = tree.iterative_uturn(is_turning)
    idx_min, idx_max = checkpoint_idxs
    state = tree.TerminationState(np.array([1.0, 2.0, 3.0, -2.0]), np.array([2.0, 4.0, 4.0, -1.0]),
                                  idx_min, idx_max)
    assert is_iterative_turning(state, momentum_sum, momentum) is expected_turning


@pytest.mark.parametrize("step, expected_idx",
                         [(0, (1, 0)), (6, (3, 2)), (7, (0, 2)), (13, (2, 2)), (15, (0, 3))])
def test_leaf_idx_to_ckpt_idx(step, expected_idx):
    assert tree._find_storage_indices(step) == expected_idx


def test_storage_indices_closed_form():
    """The popcount / trailing-ones closed form the CUDA kernels use."""
    for step in range(1, 5000):
        idx_max = bin(step >> 1).count("1")
        trailing = (~step & (step + 1)).bit_length() - 1
        assert tree._find_storage_indices(step) == (idx_max - trailing + 1, idx_max)


@pytest.mark.parametrize("num_dims", [1, 3])
def test_termination_update(num_dims):
    _, _, is_turning = hamiltonian.gaussian_metric(np.ones(1))
    new_state, update, _ = tree.iterative_uturn(is_turning)
    state = new_state(np.ones(num_dims), 4)
    ones = np.ones(num_dims)
    update(state, ones, ones, 1)
    odd = update(state, ones, ones, 5)
    np.testing.assert_array_equal(odd[0], np.zeros((4, num_dims)))
    np.testing.assert_array_equal(odd[1], np.zeros((4, num_dims)))
    even = update(state, ones, 2 * ones, 6)          # idx (3, 2): row 2 written
    np.testing.assert_array_equal(even[0][2], 2 * ones)
    assert (even.min_index, even.max_index) == (3, 2)


# ---- metrics (reference tests/test_metrics.py) -------------------------------
@pytest.mark.parametrize("imm, p, expected", [
    (1.0, 1.0, 0.5),
    (np.array([1.0]), np.array([1.0]), 0.5),
    (np.array([1.0, 1.0]), np.array([1.0, 1.0]), 1.0),
    (np.array([[1.0, 0], [0, 1.0]]), np.array([1.0, 1.0]), 1.0),
])
def test_gaussian_metric_kinetic_energy(imm, p, expected):
    _, kinetic_energy, _ = hamiltonian.gaussian_metric(imm)
    k = kinetic_energy(p)
    assert np.ndim(k) == 0 and k == expected


@pytest.mark.parametrize("imm, pl, pr, ps", [
    (1.0, 1.0, 1.0, 1.0),
    (np.ones(2), np.ones(2), np.ones(2), np.ones(2)),
    (np.eye(2), np.ones(2), np.ones(2), np.ones(2)),
])
def test_turning(imm, pl, pr, ps):
    _, _, turning = hamiltonian.gaussian_metric(imm)
    assert turning(pl, pr, ps) is True


def test_fail_wrong_mass_matrix_dimension():
    with pytest.raises(ValueError):
        hamiltonian.gaussian_metric(np.ones((2, 2, 2)))


def test_dense_momentum_has_mass_matrix_covariance():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 3))
    imm = a @ a.T + 3 * np.eye(3)

    class Z:
        def normal(self, shape):
            return np.eye(3)[0]
    gen, _, _ = hamiltonian.gaussian_metric(imm)
    # columns of mass_matrix_sqrt: S S^T = imm^{-1}
    cols = []
    for i in range(3):
        Z.normal = lambda self, shape, i=i: np.eye(3)[i]
        cols.append(gen(Z()))
    S = np.array(cols).T
    np.testing.assert_allclose(S @ S.T, np.linalg.inv(imm), rtol=1e-12)


# ---- integrators (reference tests/test_integrators.py) ------------------------
def _free_fall(q):
    return np.sum(q), np.ones_like(q)


def _harmonic(q):
    return np.sum(0.5 * np.square(q)), q


def _circular(q):
    r2 = q[0] ** 2 + q[1] ** 2
    return -1.0 / np.sqrt(r2), q / r2 ** 1.5


@pytest.mark.parametrize("potential, n_steps, q0, p0, qf, pf", [
    (_free_fall, 100, [0.0], [1.0], [0.5], [0.0]),
    (_harmonic, 100, [0.0], [1.0], [np.sin(1.0)], [np.cos(1.0)]),
    (_circular, 628, [1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.0, 1.0]),
])
def test_velocity_verlet(potential, n_steps, q0, p0, qf, pf):
    q0, p0 = np.array(q0), np.array(p0)
    _, kinetic_energy, _ = hamiltonian.gaussian_metric(np.ones(len(q0)))
    step = hamiltonian.velocity_verlet(potential, kinetic_energy)
    state = hamiltonian.new_integrator_state(potential, q0, p0)
    e0 = state.potential_energy + kinetic_energy(p0)
    integrate = tree.static_integration(step, n_steps)
    state, _ = integrate(state, 0.01)
    np.testing.assert_allclose(state.position, qf, atol=1e-2)
    np.testing.assert_allclose(state.momentum, pf, atol=1e-2)
    assert state.potential_energy + kinetic_energy(state.momentum) == pytest.approx(e0, 1e-4)


# ---- algorithms (reference tests/test_algorithms.py) --------------------------
def test_dual_averaging():
    init, update = adaptation.dual_averaging(gamma=0.5)
    state = init(0.5)
    for _ in range(100):
        state = update(2 * (state.iterates - 1), state)
    assert state.iterates_avg == pytest.approx(1.0, 1e-2)
    assert state.iterates == pytest.approx(1.0, 1e-2)


@pytest.mark.parametrize("do_compute_covariance", [True, False])
@pytest.mark.parametrize("n_dim", [0, 1, 3])
def test_welford(n_dim, do_compute_covariance):
    init, update, final = adaptation.welford_covariance(do_compute_covariance)
    state = init(n_dim)
    for i in range(10):
        state = update(float(i) if n_dim == 0 else i * np.ones(n_dim), *state)
    cov = final(state[1], state[2])
    if n_dim == 0:
        assert state[0] == 4.5 and cov == pytest.approx(55.0 / 6.0)
    else:
        np.testing.assert_allclose(state[0], 4.5 * np.ones(n_dim))
        shape = (n_dim, n_dim) if do_compute_covariance else (n_dim,)
        np.testing.assert_allclose(cov, 55.0 / 6.0 * np.ones(shape))


@pytest.mark.parametrize("do_compute_covariance", [True, False])
@pytest.mark.parametrize("num_dims", [0, 1, 3])
def test_welford_constant(num_dims, do_compute_covariance):
    init, update, final = adaptation.welford_covariance(do_compute_covariance)
    state = init(num_dims)
    for _ in range(10):
        state = update(1.0 if num_dims == 0 else np.ones(num_dims), *state)
    np.testing.assert_allclose(state[0], 1.0)
    np.testing.assert_allclose(final(state[1], state[2]), 0.0)


# ---- schedule (reference tests/test_adaptation.py) ----------------------------
@pytest.mark.parametrize("num_steps, expected_schedule", [
    (19, [(0, False)] * 19),
    (100, [(0, False)] * 15 + [(1, False)] * 74 + [(1, True)] + [(0, False)] * 10),
    (200, [(0, False)] * 75 + [(1, False)] * 24 + [(1, True)] + [(1, False)] * 49 + [(1, True)]
     + [(0, False)] * 50),
])
def test_adaptation_schedule(num_steps, expected_schedule):
    schedule = adaptation.build_schedule(num_steps)
    assert len(schedule) == num_steps and schedule == expected_schedule


def test_schedule_default_1000():
    s = adaptation.build_schedule(1000)
    ends = [i for i, (_, e) in enumerate(s) if e]
    assert ends == [99, 149, 249, 449, 949]
    assert [st for st, _ in s] == [0] * 75 + [1] * 875 + [0] * 50
