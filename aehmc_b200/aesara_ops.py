"""Aesara `Op`s over the C-ABI (the drop-in surface for an aehmc user's graph).

Import-guarded: Aesara is not installable in the build image, so this module carries no logic of its own --
each `perform` marshals NumPy inputs to the batched entry points of this package (which call libb200hmc.so
through ctypes) and writes NumPy outputs.  See INTEGRATION.md for the graph-side usage.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - exercised only where Aesara exists
    import aesara.tensor as at
    from aesara.graph.basic import Apply
    from aesara.graph.op import Op
    HAVE_AESARA = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_AESARA = False
    Op = object


def _require():
    if not HAVE_AESARA:
        raise ImportError("aehmc_b200.aesara_ops needs Aesara (>= 2.8.11), which is not installed")


class NUTSStepOp(Op):
    """(q[C,d], U[C], dU[C,d], step_size[C], imm) -> (q', p', U', dU', acceptance_probability, num_doublings,
    is_turning, is_diverging): one `nuts.new_kernel(...)` transition (reference nuts.py:56-153) of every chain."""

    __props__ = ("max_num_expansions", "divergence_threshold")

    def __init__(self, srng, logprob_fn, max_num_expansions=10, divergence_threshold=1000):
        _require()
        self.srng, self.model = srng, logprob_fn
        self.max_num_expansions, self.divergence_threshold = max_num_expansions, divergence_threshold

    def make_node(self, q, U, g, step_size, imm):
        ins = [at.as_tensor_variable(x) for x in (q, U, g, step_size, imm)]
        outs = [ins[0].type(), ins[0].type(), ins[1].type(), ins[0].type(), at.dvector(), at.ivector(),
                at.bvector(), at.bvector()]
        return Apply(self, ins, outs)

    def perform(self, node, inputs, output_storage):
        from . import nuts
        from .integrators import IntegratorState
        q, U, g, step_size, imm = inputs
        kernel = nuts.new_kernel(self.srng, self.model, self.max_num_expansions, self.divergence_threshold)
        info, _ = kernel(IntegratorState(q, None, U, g), np.asarray(step_size), imm)
        vals = (info.state.position, info.state.momentum, info.state.potential_energy,
                info.state.potential_energy_grad, info.acceptance_probability, info.num_doublings,
                info.is_turning, info.is_diverging)
        for store, v in zip(output_storage, vals):
            store[0] = v.cpu().numpy()


class HMCStepOp(Op):
    """(q, U, dU, step_size, imm) -> (q', p', U', dU', acceptance_probability, is_diverging): one
    `hmc.new_kernel(...)` transition with a static `num_integration_steps` (reference hmc.py:77-124)."""

    __props__ = ("num_integration_steps", "divergence_threshold")

    def __init__(self, srng, logprob_fn, num_integration_steps, divergence_threshold=1000):
        _require()
        self.srng, self.model = srng, logprob_fn
        self.num_integration_steps, self.divergence_threshold = num_integration_steps, divergence_threshold

    def make_node(self, q, U, g, step_size, imm):
        ins = [at.as_tensor_variable(x) for x in (q, U, g, step_size, imm)]
        outs = [ins[0].type(), ins[0].type(), ins[1].type(), ins[0].type(), at.dvector(), at.bvector()]
        return Apply(self, ins, outs)

    def perform(self, node, inputs, output_storage):
        from . import hmc
        from .integrators import IntegratorState
        q, U, g, step_size, imm = inputs
        kernel = hmc.new_kernel(self.srng, self.model, self.divergence_threshold)
        info, _ = kernel(IntegratorState(q, None, U, g), np.asarray(step_size), imm, self.num_integration_steps)
        vals = (info.state.position, info.state.momentum, info.state.potential_energy,
                info.state.potential_energy_grad, info.acceptance_probability, info.is_diverging)
        for store, v in zip(output_storage, vals):
            store[0] = v.cpu().numpy()


class PotentialAndGradOp(Op):
    """q[C,d] -> (U[C], dU[C,d]): `hmc.new_state` (reference hmc.py:16-40)."""

    __props__ = ()

    def __init__(self, logprob_fn):
        _require()
        self.model = logprob_fn

    def make_node(self, q):
        q = at.as_tensor_variable(q)
        return Apply(self, [q], [at.dvector(), q.type()])

    def perform(self, node, inputs, output_storage):
        U, g = self.model.potential_and_grad(inputs[0])
        output_storage[0][0] = U.cpu().numpy()
        output_storage[1][0] = g.cpu().numpy()
