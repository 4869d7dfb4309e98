"""Host-side helpers of bench.py: synthetic problems are deterministic, bf16 rounding matches torch, the reference arm's
sample shrinks with the number of requested steps, and every workload has a description naming its BASELINE config."""
import numpy as np
import pytest

import bench


def test_bf16_round_matches_torch():
    torch = pytest.importorskip("torch")
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32) * np.float32(3.0)
    ours = bench.bf16_round(x)
    theirs = torch.from_numpy(x).bfloat16().float().numpy()
    assert np.array_equal(ours, theirs)


def test_problems_are_deterministic_and_well_formed():
    cov, prec = bench.make_dense_problem(16)
    cov2, _ = bench.make_dense_problem(16)
    assert np.array_equal(cov, cov2) and np.allclose(cov @ prec, np.eye(16), atol=1e-9)
    X, y, imm = bench.make_logistic_problem(256, 8)
    assert X.shape == (256, 8) and set(np.unique(y)) <= {0.0, 1.0} and np.allclose(imm, 4.0 / 256)
    assert np.array_equal(bench.bf16_round(X.astype(np.float32)).astype(np.float64), X)      # bf16-representable
    a = bench.initial_positions("logistic", 4, 8, chain_offset=0)
    b = bench.initial_positions("logistic", 4, 8, chain_offset=4)
    assert a.shape == (4, 8) and not np.array_equal(a, b)


def test_reference_arm_sample_is_bounded():
    for name in bench.WORKLOADS:
        full = bench.cpu_transitions(name, 1)
        many = bench.cpu_transitions(name, 50)
        assert 3 <= many <= full
        assert name in bench.describe(name)
    assert "configs[1]" in bench.describe("c2") and "configs[2]" in bench.describe("c3") and "configs[4]" in bench.describe("c5")
