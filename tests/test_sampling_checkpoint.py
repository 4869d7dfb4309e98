"""sampling.checkpoint / restore: the wire format round-trips on a machine without a GPU (values, dtypes, Philox
coordinates); the bit-exact continuation itself is a GPU test (tests/test_gpu_storage.py)."""
import io

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_checkpoint_round_trip_cpu():
    from aehmc_b200 import sampling
    from aehmc_b200.integrators import IntegratorState
    from aehmc_b200.random import InjectedDraws, RandomStream
    rng = np.random.default_rng(0)
    q, U, g = rng.standard_normal((5, 3)), rng.standard_normal(5), rng.standard_normal((5, 3))
    state = IntegratorState(torch.from_numpy(q), None, torch.from_numpy(U), torch.from_numpy(g))
    srng = RandomStream(seed=2 ** 40 + 7, chain_offset=1234)
    srng.advance(17)
    ck = sampling.checkpoint(state, srng, 0.3, np.array([1.0, 2.0, 0.5]))
    buf = io.BytesIO()
    torch.save(ck, buf)
    buf.seek(0)
    state2, srng2, eps, imm = sampling.restore(torch.load(buf), device="cpu")
    assert (srng2.seed, srng2.chain_offset, srng2.transition) == (2 ** 40 + 7, 1234, 17)
    assert eps.dtype == torch.float64 and float(eps) == 0.3          # a Python float must not be narrowed to float32
    assert torch.equal(state2.position, state.position) and torch.equal(state2.potential_energy_grad, state.potential_energy_grad)
    assert state2.momentum is None and torch.equal(imm, torch.tensor([1.0, 2.0, 0.5], dtype=torch.float64))
    with pytest.raises(ValueError):
        sampling.restore({**ck, "version": 99}, device="cpu")
    with pytest.raises(TypeError):
        sampling.checkpoint(state, InjectedDraws.__new__(InjectedDraws), 0.3, 1.0)


def test_checkpoint_keeps_the_metric_kind_cpu():
    """window adaptation returns a [C, d] per-chain diagonal: it must come back as per_chain, never as a dense matrix."""
    from aehmc_b200 import metrics, sampling
    from aehmc_b200.integrators import IntegratorState
    from aehmc_b200.random import RandomStream
    q = torch.zeros((4, 4), dtype=torch.float64)
    state = IntegratorState(q, None, torch.zeros(4, dtype=torch.float64), q.clone())
    imm = torch.rand((4, 4), dtype=torch.float64) + 0.5            # C == d: the ambiguous case
    ck = sampling.checkpoint(state, RandomStream(1), torch.full((4,), 0.1, dtype=torch.float64), metrics.per_chain(imm))
    assert ck["inverse_mass_matrix_kind"] == "per_chain"
    _, _, eps, imm2 = sampling.restore(ck, device="cpu")
    assert isinstance(imm2, metrics.per_chain) and torch.equal(imm2.imm, imm) and eps.shape == (4,)
    for raw, kind in ((2.0, "scalar"), (np.ones(4), "diag"), (np.eye(4), "dense")):
        assert sampling.checkpoint(state, RandomStream(1), 0.1, raw)["inverse_mass_matrix_kind"] == kind
    with pytest.raises(ValueError):
        sampling.checkpoint(state, RandomStream(1), 0.1, np.ones((2, 2, 2)))
