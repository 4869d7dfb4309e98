"""CPU oracle for the aehmc trajectory hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy float64, single-chain, loop-for-loop restatement of the
algorithm defined by the reference's symbolic graphs:

    aehmc/integrators.py, metrics.py, proposals.py, termination.py,
    trajectory.py, hmc.py, nuts.py, algorithms.py, step_size.py,
    mass_matrix.py, window_adaptation.py          (paths relative to the reference)

The arithmetic of that path lives in third-party Aesara (>= 2.8.11, unpinned,
reference pyproject.toml:15-20) which is not installable here, so this is a
restatement of the graphs' published semantics, not an import of the reference.

Parity pin: the README quick-start draw 1.1034719409361107 (reference
README.md:22-55) is reproduced bit-for-bit through ``oracle.streams.StreamDraws``
(see tests/test_oracle_golden.py), together with the reference's exact unit
fixtures (tests/test_termination.py, test_metrics.py, test_adaptation.py,
test_algorithms.py, test_trajectory.py).  HMC, diagonal/dense metrics and window
adaptation have no per-transition numeric golden upstream: for those, parity is
pinned only by the reference's exact/analytic unit fixtures ("parity unpinned"
beyond them).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under aehmc_b200/ does.
"""

from . import adaptation, hamiltonian, kernels, models, streams, tree  # noqa: F401
