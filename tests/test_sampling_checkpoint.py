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
