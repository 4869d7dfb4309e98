"""Convergence diagnostics for many chains: R-hat and effective sample size.

The reference has no diagnostics of its own; its tests use ``arviz.ess`` (reference
tests/test_hmc.py:158-167), which is not installable here, and BASELINE.json's second metric is ESS/s.
This module restates the Stan / arviz estimator (Vehtari et al. 2021: multi-chain autocorrelation,
Geyer's initial monotone sequence) in two layers:

* per-chain means and autocovariances are computed on the device (``b2h_chain_autocov``) and reduced
  over chains to a few sufficient statistics; with ``torch.distributed`` initialised these are summed
  over ranks -- the only collective of the sampler (NCCL on GPUs, gloo in the CPU tests);
* the final O(lags) arithmetic runs on the host in NumPy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, backend


# ------------------------------------------------------------------------------------------------
# host arithmetic on sufficient statistics
# ------------------------------------------------------------------------------------------------
def sufficient_statistics_numpy(draws, max_lag):
    """draws [T, C, d] (NumPy) -> dict of per-dimension sums over chains (float64)."""
    x = np.asarray(draws, dtype=np.float64)
    T, Cn, d = x.shape
    mean = x.mean(0)                                            # [C, d]
    xc = x - mean
    acov = np.empty((Cn, d, max_lag + 1))
    for lag in range(max_lag + 1):
        acov[:, :, lag] = (xc[: T - lag] * xc[lag:]).sum(0) / T
    return {"n_chains": float(Cn), "n_draws": float(T), "sum_mean": mean.sum(0), "sum_mean_sq": (mean ** 2).sum(0),
            "sum_acov": acov.sum(0)}


def all_reduce_statistics(stats, device=None):
    """Sum the sufficient statistics over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    out = dict(stats)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    flat = np.concatenate([[stats["n_chains"]], stats["sum_mean"], stats["sum_mean_sq"], stats["sum_acov"].ravel()])
    t = torch.as_tensor(flat, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    flat = t.cpu().numpy()
    d = stats["sum_mean"].shape[0]
    out["n_chains"] = float(flat[0])
    out["sum_mean"] = flat[1:1 + d]
    out["sum_mean_sq"] = flat[1 + d:1 + 2 * d]
    out["sum_acov"] = flat[1 + 2 * d:].reshape(stats["sum_acov"].shape)
    return out


def rhat_from_statistics(stats):
    """Potential scale reduction (non-split, non-rank-normalised) per dimension."""
    m, n = stats["n_chains"], stats["n_draws"]
    W = stats["sum_acov"][:, 0] / m * n / (n - 1.0)
    if m < 2:
        return np.full_like(W, np.nan)
    B_over_n = (stats["sum_mean_sq"] - stats["sum_mean"] ** 2 / m) / (m - 1.0)
    var_plus = W * (n - 1.0) / n + B_over_n
    return np.sqrt(var_plus / W)


def ess_from_statistics(stats):
    """Effective sample size per dimension from chain-summed autocovariances (Stan / arviz ``_ess``:
    rho_t = 1 - (W - mean_c acov_c(t)) / var_plus, Geyer initial positive + monotone sequence)."""
    m, n = stats["n_chains"], stats["n_draws"]
    acov = stats["sum_acov"] / m                                # [d, L+1] mean over chains
    d, L1 = acov.shape
    mean_var = acov[:, 0] * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus = var_plus + (stats["sum_mean_sq"] - stats["sum_mean"] ** 2 / m) / (m - 1.0)
    out = np.empty(d)
    for j in range(d):
        if not np.isfinite(var_plus[j]) or var_plus[j] <= 0:
            out[j] = np.nan
            continue
        rho = np.zeros(L1 + 2)
        rho[0] = 1.0
        lim = L1 - 1
        get = lambda t: 1.0 - (mean_var[j] - acov[j, t]) / var_plus[j]
        if lim >= 1:
            rho[1] = get(1)
        t = 1
        rho_even, rho_odd = 1.0, rho[1]
        while t < lim - 2 and rho_even + rho_odd >= 0.0:
            rho_even, rho_odd = get(t + 1), get(t + 2)
            if rho_even + rho_odd >= 0.0:
                rho[t + 1], rho[t + 2] = rho_even, rho_odd
            t += 2
        max_t = t - 2
        if rho_even > 0:
            rho[max_t + 1] = rho_even
        t = 1
        while t <= max_t - 2:                                   # Geyer's initial monotone sequence
            if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
                rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
                rho[t + 2] = rho[t + 1]
            t += 2
        total = m * n
        tau = -1.0 + 2.0 * np.sum(rho[: max_t + 1]) + np.sum(rho[max_t + 1: max_t + 2])
        tau = max(tau, 1.0 / np.log10(total))
        out[j] = total / tau
    return out


# ------------------------------------------------------------------------------------------------
# device front end
# ------------------------------------------------------------------------------------------------
def sufficient_statistics(draws, max_lag=None, dims=None):
    """draws [T, C, d] CUDA tensor -> chain-summed statistics (means / autocovariances on the device)."""
    if not isinstance(draws, torch.Tensor) or not draws.is_cuda:
        x = np.asarray(draws.cpu() if isinstance(draws, torch.Tensor) else draws)
        if dims is not None:
            x = x[:, :, dims]
        return sufficient_statistics_numpy(x, max_lag if max_lag is not None else min(x.shape[0] - 1, 200))
    if dims is not None:
        draws = draws[:, :, dims]
    draws = draws.contiguous()
    T, Cn, d = draws.shape
    if max_lag is None:
        max_lag = min(T - 1, 200)
    lib = _lib.load()
    dev = draws.device
    mean = torch.empty((Cn, d), dtype=torch.float64, device=dev)
    acov = torch.empty((Cn, d, max_lag + 1), dtype=torch.float64, device=dev)
    _lib.check(lib.b2h_chain_autocov(backend.context(dev), backend.code(draws.dtype), backend.ptr(draws), C.c_int64(T),
                                     C.c_int64(Cn), C.c_int64(d), C.c_int32(max_lag), backend.ptr(mean), backend.ptr(acov)))
    return {"n_chains": float(Cn), "n_draws": float(T), "sum_mean": mean.sum(0).cpu().numpy(),
            "sum_mean_sq": (mean * mean).sum(0).cpu().numpy(), "sum_acov": acov.sum(0).cpu().numpy()}


def ess(draws, max_lag=None, dims=None, distributed=True, method="bulk"):
    """Effective sample size per dimension over all chains (and all ranks when distributed).
    ``method="bulk"`` (default, what ``arviz.ess`` computes by default and the reference's tests use,
    tests/test_hmc.py:158-167): chains split in halves, draws rank-normalised over the whole pool, then the
    multi-chain autocorrelation estimator.  ``method="mean"``: the same estimator on the raw draws, no split."""
    if method == "bulk":
        z = rank_normalized_split(draws, dims, distributed)
        if max_lag is None:
            max_lag = min(z.shape[0] - 1, 200)
        stats = sufficient_statistics(z, max_lag)
    elif method == "mean":
        stats = sufficient_statistics(draws, max_lag, dims)
    else:
        raise ValueError("method must be 'bulk' or 'mean'")
    if distributed:
        stats = all_reduce_statistics(stats, draws.device if isinstance(draws, torch.Tensor) and draws.is_cuda else None)
    return ess_from_statistics(stats)


def rhat(draws, dims=None, distributed=True, method="rank"):
    """Potential scale reduction per dimension.  ``method="rank"`` (default, ``arviz.rhat``'s default; Vehtari et al.
    2021): the larger of the split-R-hat of the rank-normalised draws and of the rank-normalised folded draws
    |x - median|.  ``method="split"``: split chains, raw draws.  ``method="identity"``: raw chains."""
    cuda_dev = draws.device if isinstance(draws, torch.Tensor) and draws.is_cuda else None

    def from_draws(x):
        stats = sufficient_statistics(x, 1)
        if distributed:
            stats = all_reduce_statistics(stats, cuda_dev)
        return rhat_from_statistics(stats)

    if method == "rank":
        bulk = from_draws(rank_normalized_split(draws, dims, distributed))
        tail = from_draws(rank_normalized_split(draws, dims, distributed, fold=True))
        return np.maximum(bulk, tail)
    if method == "split":
        x = _as_tensor(draws)
        if dims is not None:
            x = x[:, :, dims]
        return from_draws(split_chains(x))
    if method == "identity":
        stats = sufficient_statistics(draws, 1, dims)
        if distributed:
            stats = all_reduce_statistics(stats, cuda_dev)
        return rhat_from_statistics(stats)
    raise ValueError("method must be 'rank', 'split' or 'identity'")


# ------------------------------------------------------------------------------------------------
# split chains + rank normalisation (arviz _split_chains / _z_scale), on the device of the draws
# ------------------------------------------------------------------------------------------------
def _as_tensor(draws):
    return draws if isinstance(draws, torch.Tensor) else torch.as_tensor(np.asarray(draws))


def split_chains(x):
    """[T, C, d] -> [T // 2, 2 C, d]: first and last half of every chain as separate chains."""
    half = x.shape[0] // 2
    return torch.cat([x[:half], x[x.shape[0] - half:]], dim=1)


def _gather_sorted(v):
    """Sorted values of every rank ([n_r] each).  Without a process group: just the local ones."""
    import torch.distributed as dist
    s = torch.sort(v).values
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [s]
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64, device=v.device)
    counts[dist.get_rank()] = s.numel()
    dist.all_reduce(counts)
    nmax = int(counts.max())
    pad = s if s.numel() == nmax else torch.cat([s, s.new_full((nmax - s.numel(),), float("inf"))])
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)                       # the "gather of draws" of BASELINE.json configs[4]
    return [p[:int(n)] for p, n in zip(parts, counts.tolist())]


def rank_normalized_split(draws, dims=None, distributed=True, fold=False):
    """z-scores of the split chains: z = Phi^-1((rank - 3/8) / (S + 1/4)), average ranks over the pool of ALL draws of
    all chains (of all ranks when ``distributed``: every rank sorts its own values, the sorted runs are all-gathered,
    and each rank ranks its own draws against all runs).  ``fold``: ranks of |x - median| (tail R-hat).
    Returns [T // 2, 2 C, d'] float64 on the device of ``draws``."""
    x = _as_tensor(draws)
    if dims is not None:
        x = x[:, :, dims]
    x = split_chains(x).to(torch.float64)
    out = torch.empty_like(x)
    for j in range(x.shape[2]):
        v = x[:, :, j].reshape(-1).contiguous()
        runs = _gather_sorted(v) if distributed else [torch.sort(v).values]
        if fold:
            allv = torch.sort(torch.cat(runs)).values
            n = allv.numel()
            med = allv[n // 2] if n % 2 else 0.5 * (allv[n // 2 - 1] + allv[n // 2])
            v = (v - med).abs()
            runs = [torch.sort((r - med).abs()).values for r in runs]
        total = sum(int(r.numel()) for r in runs)
        less = torch.zeros_like(v)
        leq = torch.zeros_like(v)
        for r in runs:
            less += torch.searchsorted(r, v, right=False).to(torch.float64)
            leq += torch.searchsorted(r, v, right=True).to(torch.float64)
        rank = 0.5 * (less + leq + 1.0)                # average rank, 1-based (scipy.stats.rankdata "average")
        out[:, :, j] = torch.special.ndtri((rank - 0.375) / (total + 0.25)).reshape(x.shape[0], x.shape[1])
    return out


def gather_draws(draws, dims=None, thin=1):
    """All-gather (thinned, dimension-selected) draws over the ranks: [T, C_local, d] -> [T', C_total, len(dims)] on
    every rank, chains in global order (BASELINE.json configs[4]: "NCCL gather of draws for R-hat / ESS").  The
    sufficient-statistics all-reduce of :func:`ess` / :func:`rhat` moves (2 + lags) * d doubles instead and is what the
    benchmark uses; the gather is for consumers that need the draws themselves (plots, rank-normalised diagnostics).
    A full gather at the c5 size is 1 GB per stored draw: select dimensions and thin."""
    import torch.distributed as dist
    x = draws if isinstance(draws, torch.Tensor) else torch.as_tensor(np.asarray(draws))
    x = x[::max(int(thin), 1)]
    if dims is not None:
        x = x[:, :, dims]
    x = x.contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64, device=x.device)
    counts[dist.get_rank()] = x.shape[1]
    dist.all_reduce(counts)
    cmax = int(counts.max())
    pad = x if x.shape[1] == cmax else torch.cat(
        [x, x.new_zeros((x.shape[0], cmax - x.shape[1], x.shape[2]))], dim=1)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous())
    return torch.cat([p[:, :int(n)] for p, n in zip(parts, counts.tolist())], dim=1)


def ess_bulk_single_chain(x):
    """arviz.ess(method="bulk") for ONE chain x[T] as the reference's tests call it: split the chain in two,
    rank-normalise over all draws, then the estimator above."""
    from scipy import stats as sstats
    x = np.asarray(x, dtype=np.float64)
    half = x.shape[0] // 2
    n = 2 * half
    halves = np.stack([x[:half], x[x.shape[0] - half:]])        # arviz _split_chains: an odd middle draw is dropped
    ranks = sstats.rankdata(halves.ravel(), method="average").reshape(halves.shape)
    z = sstats.norm.ppf((ranks - 0.375) / (halves.size + 0.25))
    st = sufficient_statistics_numpy(z.T[:, :, None], n // 2 - 1)
    return float(ess_from_statistics(st)[0])


def shard_chains(num_chains, rank=None, world_size=None):
    """Contiguous block of global chain ids owned by ``rank``: (offset, count).  Chains are independent in
    the reference (one chain per compiled function), so sharding needs no data-path collective; Philox is
    keyed by the GLOBAL chain id, so results do not depend on the number of GPUs."""
    import torch.distributed as dist
    if rank is None or world_size is None:
        if dist.is_available() and dist.is_initialized():
            rank, world_size = dist.get_rank(), dist.get_world_size()
        else:
            rank, world_size = 0, 1
    base, rem = divmod(int(num_chains), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count
