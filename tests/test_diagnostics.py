"""ESS / R-hat arithmetic (host part) and the N>1 reduction path on the gloo backend."""
import os
import sys

import numpy as np
import pytest


def _load():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import aehmc_b200.diagnostics as d
    return d


def _ar1(rng, phi, T, C):
    x = np.zeros((T, C, 1))
    e = rng.standard_normal((T, C, 1)) * np.sqrt(1 - phi ** 2)
    x[0] = rng.standard_normal((C, 1))
    for t in range(1, T):
        x[t] = phi * x[t - 1] + e[t]
    return x


@pytest.mark.parametrize("phi", [0.0, 0.5, 0.9, -0.5])
def test_ess_matches_ar1_theory(phi):
    diag = _load()
    rng = np.random.default_rng(0)
    T, C = 2000, 16
    x = _ar1(rng, phi, T, C)
    stats = diag.sufficient_statistics_numpy(x, 300)
    ess = diag.ess_from_statistics(stats)[0]
    theory = T * C * (1 - phi) / (1 + phi)
    assert ess == pytest.approx(theory, rel=0.15)
    assert diag.rhat_from_statistics(stats)[0] == pytest.approx(1.0, abs=0.01)


def test_rhat_detects_disagreeing_chains():
    diag = _load()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((500, 8, 2))
    x[:, :4, 1] += 3.0
    r = diag.rhat_from_statistics(diag.sufficient_statistics_numpy(x, 1))
    assert r[0] < 1.02 and r[1] > 1.5


def test_bulk_ess_single_chain_iid():
    diag = _load()
    x = np.random.default_rng(2).standard_normal(4000)
    assert diag.ess_bulk_single_chain(x) == pytest.approx(4000, rel=0.15)


def _bulk_reference(x):
    """arviz.ess(method="bulk") / arviz.rhat(method="rank") restated with scipy (independent of the torch code):
    _split_chains, _z_scale with scipy.stats.rankdata, then the estimator on the z-scores."""
    from scipy import stats as sstats
    diag = _load()
    T, C, d = x.shape
    half = T // 2
    split = np.concatenate([x[:half], x[T - half:]], axis=1)              # [half, 2C, d]
    ess, rhat = np.empty(d), np.empty(d)
    for j in range(d):
        def zscale(a):
            r = sstats.rankdata(a.ravel(), method="average").reshape(a.shape)
            return sstats.norm.ppf((r - 0.375) / (a.size + 0.25))
        z = zscale(split[:, :, j])
        ess[j] = diag.ess_from_statistics(diag.sufficient_statistics_numpy(z[:, :, None], min(half - 1, 200)))[0]
        zf = zscale(np.abs(split[:, :, j] - np.median(split[:, :, j])))
        rb = diag.rhat_from_statistics(diag.sufficient_statistics_numpy(z[:, :, None], 1))[0]
        rt = diag.rhat_from_statistics(diag.sufficient_statistics_numpy(zf[:, :, None], 1))[0]
        rhat[j] = max(rb, rt)
    return ess, rhat


def test_bulk_ess_and_rank_rhat_match_the_scipy_restatement():
    import torch
    diag = _load()
    rng = np.random.default_rng(3)
    x = np.concatenate([_ar1(rng, 0.7, 401, 6), np.exp(_ar1(rng, 0.2, 401, 6)), np.round(_ar1(rng, 0.0, 401, 6), 1)], axis=2)
    ess_ref, rhat_ref = _bulk_reference(x)                               # odd T, heavy tail, ties
    ess = diag.ess(torch.from_numpy(x), distributed=False)
    rhat = diag.rhat(torch.from_numpy(x), distributed=False)
    np.testing.assert_allclose(ess, ess_ref, rtol=1e-9)
    np.testing.assert_allclose(rhat, rhat_ref, rtol=1e-9)
    one = x[:, :1, :1]
    assert diag.ess(torch.from_numpy(one), distributed=False)[0] == pytest.approx(diag.ess_bulk_single_chain(one[:, 0, 0]), rel=1e-9)


def test_rank_rhat_sees_what_plain_rhat_misses():
    """Chains with the same mean but different scales: only the folded (tail) rank-normalised split R-hat reacts;
    a chain that drifts: only the SPLIT R-hat reacts."""
    import torch
    diag = _load()
    rng = np.random.default_rng(4)
    x = rng.standard_normal((400, 8, 2))
    x[:, :4, 0] *= 4.0
    x[:, :, 1] += np.linspace(-1.5, 1.5, 400)[:, None]
    t = torch.from_numpy(x)
    plain = diag.rhat(t, distributed=False, method="identity")
    rank = diag.rhat(t, distributed=False)
    assert plain[0] < 1.01 and rank[0] > 1.1
    assert plain[1] < 1.01 and rank[1] > 1.15 and diag.rhat(t, distributed=False, method="split")[1] > 1.15


def _rank_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aehmc_b200.diagnostics as diag
    x = np.exp(_ar1(np.random.default_rng(11), 0.5, 200, 7))
    off, cnt = diag.shard_chains(7)                                       # 4 + 3 chains
    local = torch.from_numpy(x[:, off:off + cnt])
    q.put((rank, float(diag.ess(local)[0]), float(diag.rhat(local)[0])))
    dist.destroy_process_group()


def test_rank_normalised_diagnostics_over_ranks_gloo():
    """Global ranks from all-gathered sorted runs: sharded chains give the single-process bulk-ESS / rank R-hat."""
    import torch
    import torch.multiprocessing as mp
    diag = _load()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    x = torch.from_numpy(np.exp(_ar1(np.random.default_rng(11), 0.5, 200, 7)))
    ess, rhat = diag.ess(x, distributed=False)[0], diag.rhat(x, distributed=False)[0]
    for r in res:
        assert r[1] == pytest.approx(ess, rel=1e-10) and r[2] == pytest.approx(rhat, rel=1e-10)


def test_shard_chains_covers_everything():
    diag = _load()
    for n, w in ((10, 3), (4096, 8), (7, 8), (1 << 20, 8)):
        blocks = [diag.shard_chains(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (o1, c1), (o2, _) in zip(blocks, blocks[1:]):
            assert o1 + c1 == o2


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aehmc_b200.diagnostics as diag
    rng = np.random.default_rng(5)
    x = _ar1(rng, 0.6, 800, 12)                    # the same global draws on every rank
    off, cnt = diag.shard_chains(12)
    local = diag.sufficient_statistics_numpy(x[:, off:off + cnt], 150)
    merged = diag.all_reduce_statistics(local)
    import torch
    off7, cnt7 = diag.shard_chains(7)              # uneven shards: 4 + 3 chains
    y = _ar1(np.random.default_rng(9), 0.3, 20, 7)
    gathered = diag.gather_draws(torch.from_numpy(y[:, off7:off7 + cnt7]), dims=[0], thin=2)
    ok = bool(torch.equal(gathered, torch.from_numpy(y[::2, :, :1])))
    q.put((rank, float(diag.ess_from_statistics(merged)[0]), float(diag.rhat_from_statistics(merged)[0]), off, cnt, ok))
    dist.destroy_process_group()


def test_sharded_statistics_reduce_to_the_global_answer_gloo():
    """world_size = 2 on gloo: each rank holds a shard of the chains; the all-reduced sufficient statistics
    give the same ESS / R-hat as the single-process computation over all chains."""
    import torch.multiprocessing as mp
    diag = _load()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    x = _ar1(np.random.default_rng(5), 0.6, 800, 12)
    full = diag.sufficient_statistics_numpy(x, 150)
    ess, rhat = diag.ess_from_statistics(full)[0], diag.rhat_from_statistics(full)[0]
    assert [r[3:5] for r in res] == [(0, 6), (6, 6)]
    assert all(r[5] for r in res)                  # gather_draws: global chain order, uneven shards, thinned
    for r in res:
        assert r[1] == pytest.approx(ess, rel=1e-12) and r[2] == pytest.approx(rhat, rel=1e-12)
