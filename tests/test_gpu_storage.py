"""Thinned draw storage and the checkpoint / resume format: a thinned run stores exactly every
k-th draw of the full run, and a run continued from a saved checkpoint equals the uninterrupted run bit for bit
(Philox is keyed by seed, global chain id and transition)."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ab(cuda_device):
    import aehmc_b200
    return aehmc_b200


def _setup(ab, split):
    rng = np.random.default_rng(3)
    d, Cn = 12, 96
    if split:                      # dense metric -> split engine
        A = rng.standard_normal((d, d))
        cov = A @ A.T / d + 0.2 * np.eye(d)
        model = ab.models.CorrelatedGaussian(np.zeros(d), np.linalg.inv(cov))
        imm = cov
    else:                          # fused persistent kernel
        model = ab.models.IIDGaussian(np.zeros(d), np.exp(0.3 * rng.standard_normal(d)))
        imm = np.ones(d)
    q0 = rng.standard_normal((Cn, d))
    return model, imm, q0


@pytest.mark.parametrize("split", [False, True])
def test_thinned_storage_is_a_subsequence(ab, split):
    model, imm, q0 = _setup(ab, split)
    n, thin = 11, 3
    full = ab.sampling.sample(ab.nuts.new_kernel(ab.RandomStream(seed=5), model), ab.nuts.new_state(q0, model), 0.3, imm, n)
    thinned = ab.sampling.sample(ab.nuts.new_kernel(ab.RandomStream(seed=5), model), ab.nuts.new_state(q0, model), 0.3, imm,
                                 n, thin=thin)
    assert thinned[1].shape[0] == (n + thin - 1) // thin
    assert torch.equal(thinned[1], full[1][::thin])
    assert torch.equal(thinned[2], full[2][::thin])
    assert torch.equal(thinned[0].state.position, full[0].state.position)


@pytest.mark.parametrize("split", [False, True])
def test_checkpoint_resume_is_bit_exact(ab, split):
    model, imm, q0 = _setup(ab, split)
    srng = ab.RandomStream(seed=9, chain_offset=1000)
    kernel = ab.nuts.new_kernel(srng, model)
    whole = ab.sampling.sample(kernel, ab.nuts.new_state(q0, model), 0.3, imm, 6)

    srng1 = ab.RandomStream(seed=9, chain_offset=1000)
    first = ab.sampling.sample(ab.nuts.new_kernel(srng1, model), ab.nuts.new_state(q0, model), 0.3, imm, 3)
    buf = io.BytesIO()
    torch.save(ab.sampling.checkpoint(first[0].state, srng1, 0.3, imm), buf)       # the wire format
    buf.seek(0)
    state, srng2, eps, imm2 = ab.sampling.restore(torch.load(buf), device="cuda:0")
    assert srng2.transition == 3 and srng2.chain_offset == 1000
    second = ab.sampling.sample(ab.nuts.new_kernel(srng2, model), state, eps, imm2, 3)
    assert torch.equal(torch.cat([first[1], second[1]]), whole[1])
    assert torch.equal(second[0].state.position, whole[0].state.position)


def test_checkpoint_after_window_adaptation_is_bit_exact(ab):
    """The README workflow: warm up, checkpoint (per-chain step sizes + per-chain diagonal inverse mass matrix),
    restore, keep sampling -- equal to sampling straight on, bit for bit."""
    rng = np.random.default_rng(3)
    Cn, d = 64, 10
    model = ab.models.NealFunnel(d)
    q0 = rng.standard_normal((Cn, d))
    srng = ab.RandomStream(seed=21)
    kernel = ab.nuts.new_kernel(srng, model)
    state, (eps, imm), _ = ab.window_adaptation.run(kernel, ab.nuts.new_state(q0, model), 60)
    assert imm.shape == (Cn, d)
    buf = io.BytesIO()
    torch.save(ab.sampling.checkpoint(state, srng, eps, ab.metrics.per_chain(imm)), buf)
    buf.seek(0)
    whole = ab.sampling.sample(kernel, state, eps, ab.metrics.per_chain(imm), 5)
    state2, srng2, eps2, imm2 = ab.sampling.restore(torch.load(buf, weights_only=False), device="cuda:0")
    assert isinstance(imm2, ab.metrics.per_chain) and srng2.transition == 60
    again = ab.sampling.sample(ab.nuts.new_kernel(srng2, model), state2, eps2, imm2, 5)
    assert torch.equal(again[1], whole[1]) and torch.equal(again[0].state.position, whole[0].state.position)
