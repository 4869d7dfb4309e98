"""Shared parity helpers: run the oracle chain by chain on injected draws and compare
with a batched implementation (the CUDA path, or the host simulation of engine.cuh)."""
import numpy as np

from oracle import adaptation, kernels, models, streams
from oracle.hamiltonian import IntegratorState


def random_draws(rng, C, T, d, maxd=10):
    return {
        "z": rng.standard_normal((C, T, d)),
        "u_dir": rng.random((C, T, maxd)),
        "u_biased": rng.random((C, T, maxd)),
        "u_uniform": rng.random((C, T, (1 << maxd) - 1)),
        "u_accept": rng.random((C, T)),
    }


def chain_draws(draws, c):
    return streams.InjectedDraws(draws["z"][c], draws["u_dir"][c], draws["u_biased"][c],
                                 draws["u_uniform"][c], draws["u_accept"][c])


def oracle_nuts(model, q0, eps, imm, draws, n_transitions, maxd=10, div_thr=1000.0, schedule_steps=0,
                target=0.8, init_step_size=1.0, per_chain_imm=False, exact_doubling=False):
    """Per-chain oracle run.  imm: 0-d / [d] / [C,d] / [d,d].  Returns dict of stacked outputs."""
    q0 = np.asarray(q0, dtype=np.float64)
    C, d = q0.shape
    eps = np.broadcast_to(np.asarray(eps, dtype=np.float64), (C,))
    imm = np.asarray(imm, dtype=np.float64)
    if imm.ndim == 0:          # the engine's scalar metric on a d-vector is imm * I
        imm = np.full(d, float(imm))
    out = {k: [] for k in ("q", "p", "U", "g", "acceptance_probability", "num_doublings", "is_turning",
                           "is_diverging", "n_leapfrog", "draws", "eps", "imm", "hist")}
    for c in range(C):
        srng = chain_draws(draws, c)
        kernel = kernels.nuts_new_kernel(srng, model, maxd, div_thr, exact_doubling)
        state = kernels.new_state(q0[c].copy(), model)
        imm_c = imm[c] if per_chain_imm else imm
        pos, hist = [], []
        eps_c = float(eps[c])
        if schedule_steps > 0:
            trace = []
            with np.errstate(all="ignore"):
                state, (eps_c, imm_c), _ = adaptation.run(
                    kernel, state, schedule_steps, initial_step_size=init_step_size,
                    target_acceptance_rate=target, trace=trace)
            info, extras = trace[-1][0], trace[-1][1]
            pos = [t[0].state.position for t in trace]
            hist = [(t[0].num_doublings, t[1]["n_leapfrog"]) for t in trace]
            extra_t = n_transitions - schedule_steps
        else:
            extra_t = n_transitions
        for _ in range(extra_t):
            with np.errstate(all="ignore"):
                info, extras = kernel(state, eps_c, imm_c)
            pos.append(info.state.position)
            hist.append((info.num_doublings, extras["n_leapfrog"]))
            state = IntegratorState(info.state.position, None, info.state.potential_energy,
                                    info.state.potential_energy_grad)
        out["q"].append(info.state.position); out["p"].append(info.state.momentum)
        out["U"].append(info.state.potential_energy); out["g"].append(info.state.potential_energy_grad)
        out["acceptance_probability"].append(info.acceptance_probability)
        out["num_doublings"].append(info.num_doublings); out["is_turning"].append(info.is_turning)
        out["is_diverging"].append(info.is_diverging); out["n_leapfrog"].append(extras["n_leapfrog"])
        out["draws"].append(np.array(pos)); out["eps"].append(eps_c)
        out["imm"].append(np.array(np.broadcast_to(np.asarray(imm_c, dtype=np.float64), (d,)))
                          if np.ndim(imm_c) < 2 else np.asarray(imm_c))
        out["hist"].append(hist)
    res = {k: np.array([np.asarray(x, dtype=np.float64) for x in v]) for k, v in out.items() if k != "hist"}
    res["draws"] = np.transpose(res["draws"].reshape(C, -1, d), (1, 0, 2))       # [T, C, d]
    res["hist"] = out["hist"]
    return res


def oracle_hmc(model, q0, eps, imm, draws, n_transitions, L, div_thr=1000.0, per_chain_imm=False):
    q0 = np.asarray(q0, dtype=np.float64)
    C, d = q0.shape
    eps = np.broadcast_to(np.asarray(eps, dtype=np.float64), (C,))
    imm = np.asarray(imm, dtype=np.float64)
    if imm.ndim == 0:
        imm = np.full(d, float(imm))
    out = {k: [] for k in ("q", "p", "U", "g", "acceptance_probability", "is_diverging", "draws")}
    for c in range(C):
        srng = chain_draws(draws, c)
        kernel = kernels.hmc_new_kernel(srng, model, div_thr)
        state = kernels.new_state(q0[c].copy(), model)
        imm_c = imm[c] if per_chain_imm else imm
        pos = []
        for _ in range(n_transitions):
            with np.errstate(all="ignore"):
                info, _ = kernel(state, float(eps[c]), imm_c, L)
            pos.append(info.state.position)
            state = info.state._replace(momentum=None)
        out["q"].append(info.state.position); out["p"].append(info.state.momentum)
        out["U"].append(info.state.potential_energy); out["g"].append(info.state.potential_energy_grad)
        out["acceptance_probability"].append(info.acceptance_probability)
        out["is_diverging"].append(info.is_diverging); out["draws"].append(np.array(pos))
    res = {k: np.array([np.asarray(x, dtype=np.float64) for x in v]) for k, v in out.items()}
    res["draws"] = np.transpose(res["draws"].reshape(C, -1, d), (1, 0, 2))
    return res


def assert_nuts_parity(got, ref, rtol=1e-10, atol=1e-12, what=""):
    """Integers and flags bit-exact; positions / energies within tolerance (BASELINE.json north_star)."""
    np.testing.assert_array_equal(np.asarray(got["num_doublings"]), ref["num_doublings"], err_msg=what + " num_doublings")
    np.testing.assert_array_equal(np.asarray(got["n_leapfrog"]), ref["n_leapfrog"], err_msg=what + " n_leapfrog")
    np.testing.assert_array_equal(np.asarray(got["is_turning"]).astype(bool), ref["is_turning"].astype(bool), err_msg=what + " is_turning")
    np.testing.assert_array_equal(np.asarray(got["is_diverging"]).astype(bool), ref["is_diverging"].astype(bool), err_msg=what + " is_diverging")
    for k in ("q", "p", "g", "U", "acceptance_probability"):
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), ref[k], rtol=rtol, atol=atol, err_msg=what + " " + k)
