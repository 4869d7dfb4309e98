"""Pooled (cross-chain) adaptation statistics: the cross-rank merge on gloo (world_size 2) against the reference's
own Welford recurrence over all values (oracle/adaptation.py)."""
import os

import numpy as np
import pytest

from oracle import adaptation as o_adapt

torch = pytest.importorskip("torch")


def _welford(values, full):
    init, update, _ = o_adapt.welford_covariance(full)
    mean, m2, n = init(values.shape[1])
    for v in values:
        mean, m2, n = update(v, mean, m2, n)
    return n, mean, m2


@pytest.mark.parametrize("full", [False, True])
def test_merge_is_the_welford_recurrence(full):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((37, 5)) * np.array([1.0, 3.0, 0.2, 10.0, 1.0]) + np.array([0.0, 100.0, -5.0, 0.0, 1e3])
    n, mean, m2 = _welford(x, full)
    a, b = _welford(x[:11], full), _welford(x[11:], full)
    n2, mean2, m22 = o_adapt.merge_welford(*a, *b)
    assert n2 == n
    np.testing.assert_allclose(mean2, mean, rtol=1e-13)
    np.testing.assert_allclose(m22, m2, rtol=1e-11, atol=1e-9)
    from aehmc_b200.mass_matrix import merge_welford
    t = lambda s: (s[0], torch.from_numpy(np.asarray(s[1])), torch.from_numpy(np.asarray(s[2])))
    n3, mean3, m23 = merge_welford(*t(a), *t(b))
    np.testing.assert_allclose(mean3.numpy(), mean2, rtol=1e-15)
    np.testing.assert_allclose(m23.numpy(), m22, rtol=1e-14)


def _worker(rank, world, port, q, full):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aehmc_b200.mass_matrix import PooledWelford
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 10, 4)) * np.array([1.0, 2.0, 0.5, 4.0]) + np.array([0.0, 5.0, -1.0, 50.0])   # [T, C, d]
    lo, hi = (0, 6) if rank == 0 else (6, 10)                      # uneven shards of the chains
    n, mean, m2 = _welford(x[:, lo:hi].reshape(-1, 4), full)
    pool = PooledWelford(4, full, device="cpu")
    pool.n, pool.mean, pool.m2 = n, torch.from_numpy(mean), torch.from_numpy(np.asarray(m2))
    pool.all_reduce()
    q.put((rank, pool.n, pool.mean.numpy(), pool.m2.numpy(), pool.final().numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("full", [False, True])
def test_pooled_welford_all_reduce_gloo(full):
    """world_size = 2 on gloo: each rank holds the Welford state of its chains; after all_reduce every rank holds the
    state -- and the inverse mass matrix -- of the single-process computation over all chains."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (1 if full else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, full)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 10, 4)) * np.array([1.0, 2.0, 0.5, 4.0]) + np.array([0.0, 5.0, -1.0, 50.0])
    n, mean, m2 = _welford(x.reshape(-1, 4), full)
    _, _, final = o_adapt.covariance_adaptation(full)
    imm = final((mean, m2, n))
    for r in res:
        assert r[1] == n == 30
        np.testing.assert_allclose(r[2], mean, rtol=1e-13)
        np.testing.assert_allclose(r[3], m2, rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(r[4], imm, rtol=1e-11, atol=1e-12)
    assert np.array_equal(res[0][3], res[1][3])                   # identical bits on every rank
