"""Random draws for the oracle (test infrastructure).

Two interchangeable providers:

* ``StreamDraws`` reproduces the RNG topology of ``aesara.tensor.random.utils.
  RandomStream`` as used by the reference: every ``srng.<dist>()`` *call site*
  owns a ``numpy.random.default_rng`` seeded with the next ``SeedSequence(seed)
  .spawn(1)`` child in graph-construction order.  NUTS call sites, in order:
  momentum normal (reference nuts.py:113 -> metrics.py:66), direction Bernoulli
  (trajectory.py:516), uniform progressive sampling (proposals.py:99), biased
  progressive sampling (proposals.py:131).  HMC: momentum (hmc.py:122), accept
  (hmc.py:194).  ``bernoulli(p)`` is ``Generator.binomial(1, p)``.

* ``InjectedDraws`` is the validation-mode provider shared with the CUDA path:
  index-addressed arrays of standard normals and uniforms, with the Bernoulli
  decision rule that ``Generator.binomial(1, p)`` applies to its one uniform
  (``bernoulli_from_uniform``).
"""
from __future__ import annotations

import numpy as np


# Optional decision log (tests only): when set to a list, every Bernoulli decision appends ("bernoulli", margin) with
# margin = distance of the uniform from the decision threshold, and every U-turn test appends ("uturn", cosine): how
# far each decision was from flipping.  Used to show that float32 tree mismatches are near-ties.
MARGIN_LOG = None


def bernoulli_from_uniform(u: float, p: float) -> bool:
    """Decision NumPy's ``Generator.binomial(1, p)`` takes given its uniform ``u``.

    p <= 0.5 uses the inversion branch (success iff u > 1-p); p > 0.5 uses the
    mirrored branch (success iff u <= p).  NaN p never accepts.
    """
    if MARGIN_LOG is not None and p == p:
        MARGIN_LOG.append(("bernoulli", abs(u - (1.0 - p)) if p <= 0.5 else abs(u - p)))
    if p <= 0.5:
        return bool(u > 1.0 - p)
    return bool(u <= p)


def uniform_slot(expansion: int, step: int) -> int:
    """Flat slot of the uniform-sampling draw of sub-tree step ``step`` (>=1) in
    expansion ``expansion``: expansions before it hold 2**k slots each."""
    return (1 << expansion) - 1 + (step - 1)


class StreamDraws:
    def __init__(self, seed: int, kind: str = "nuts", first_child: int = 0):
        children = np.random.SeedSequence(seed).spawn(first_child + 4)[first_child:]
        gens = [np.random.default_rng(c) for c in children]
        self.momentum_rng = gens[0]
        if kind == "nuts":
            self.direction_rng, self.uniform_rng, self.biased_rng = gens[1:4]
        elif kind == "hmc":
            self.accept_rng = gens[1]
        else:
            raise ValueError(kind)

    def begin_transition(self):
        pass

    def normal(self, shape):
        return self.momentum_rng.normal(0.0, 1.0, size=shape)

    def direction(self, expansion):
        return bool(self.direction_rng.binomial(1, 0.5))

    def uniform_accept(self, expansion, step, p):
        return bool(self.uniform_rng.binomial(1, p))

    def biased_accept(self, expansion, p):
        return bool(self.biased_rng.binomial(1, p))

    def hmc_accept(self, p):
        return bool(self.accept_rng.binomial(1, p))


class InjectedDraws:
    """Index-addressed draws for ONE chain, ``n_transitions`` transitions.

    z          [T, d]            standard normals for the momentum
    u_dir      [T, max_exp]      direction uniforms   (go right iff u > 0.5)
    u_biased   [T, max_exp]      biased-sampling uniforms
    u_uniform  [T, 2**max_exp-1] uniform-sampling uniforms, slot = uniform_slot(k, s)
    u_accept   [T]               HMC accept uniforms
    """

    def __init__(self, z, u_dir=None, u_biased=None, u_uniform=None, u_accept=None):
        self.z = np.asarray(z, dtype=np.float64)
        self.u_dir = None if u_dir is None else np.asarray(u_dir, dtype=np.float64)
        self.u_biased = None if u_biased is None else np.asarray(u_biased, dtype=np.float64)
        self.u_uniform = None if u_uniform is None else np.asarray(u_uniform, dtype=np.float64)
        self.u_accept = None if u_accept is None else np.asarray(u_accept, dtype=np.float64)
        self.t = -1

    @classmethod
    def random(cls, rng, n_transitions, shape, max_num_expansions=10):
        d = int(np.prod(shape)) if shape != () else 1
        return cls(
            rng.standard_normal((n_transitions, d)),
            rng.random((n_transitions, max_num_expansions)),
            rng.random((n_transitions, max_num_expansions)),
            rng.random((n_transitions, (1 << max_num_expansions) - 1)),
            rng.random((n_transitions,)),
        )

    def begin_transition(self):
        self.t += 1

    def normal(self, shape):
        z = self.z[self.t]
        return z[0] if shape == () else z.reshape(shape)

    def direction(self, expansion):
        return bernoulli_from_uniform(self.u_dir[self.t, expansion], 0.5)

    def uniform_accept(self, expansion, step, p):
        return bernoulli_from_uniform(self.u_uniform[self.t, uniform_slot(expansion, step)], p)

    def biased_accept(self, expansion, p):
        return bernoulli_from_uniform(self.u_biased[self.t, expansion], p)

    def hmc_accept(self, p):
        return bernoulli_from_uniform(self.u_accept[self.t], p)


class RecordingStreamDraws(StreamDraws):
    """``StreamDraws`` that also writes every draw it consumes into injected-layout
    arrays (one transition per ``begin_transition``), so that a transition driven by
    the reference's RNG topology can be replayed index-addressed by the CUDA path.
    Each Bernoulli takes its uniform with ``Generator.random()`` -- the same double
    ``binomial(1, p)`` would consume -- and applies ``bernoulli_from_uniform``;
    p == 0 consumes nothing, like NumPy."""

    def __init__(self, seed, kind="nuts", first_child=0, max_num_expansions=10):
        super().__init__(seed, kind, first_child)
        self.max_num_expansions = max_num_expansions
        self.rows = []

    def begin_transition(self):
        m = self.max_num_expansions
        self.rows.append({
            "z": None, "u_dir": np.full(m, 0.25), "u_biased": np.full(m, 0.25),
            "u_uniform": np.full((1 << m) - 1, 0.25), "u_accept": 0.25,
        })

    def _bern(self, rng, p):
        if p == 0.0:
            return 0.25, False
        u = rng.random()
        return u, bernoulli_from_uniform(u, p)

    def normal(self, shape):
        z = super().normal(shape)
        self.rows[-1]["z"] = np.atleast_1d(np.asarray(z, dtype=np.float64)).ravel()
        return z

    def direction(self, expansion):
        u, ok = self._bern(self.direction_rng, 0.5)
        self.rows[-1]["u_dir"][expansion] = u
        return ok

    def uniform_accept(self, expansion, step, p):
        u, ok = self._bern(self.uniform_rng, p)
        self.rows[-1]["u_uniform"][uniform_slot(expansion, step)] = u
        return ok

    def biased_accept(self, expansion, p):
        u, ok = self._bern(self.biased_rng, p)
        self.rows[-1]["u_biased"][expansion] = u
        return ok

    def hmc_accept(self, p):
        u, ok = self._bern(self.accept_rng, p)
        self.rows[-1]["u_accept"] = u
        return ok

    def injected(self):
        """Arrays [T, ...] in the InjectedDraws / b2h_rng layout."""
        return {
            "z": np.stack([r["z"] for r in self.rows]),
            "u_dir": np.stack([r["u_dir"] for r in self.rows]),
            "u_biased": np.stack([r["u_biased"] for r in self.rows]),
            "u_uniform": np.stack([r["u_uniform"] for r in self.rows]),
            "u_accept": np.array([r["u_accept"] for r in self.rows]),
        }
