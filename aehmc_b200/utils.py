"""Structured parameters <-> the flat ``[chains, dim]`` position of the sampler.

Counterpart of the reference's ``RaveledParamsMap`` (reference utils.py:22-74): the reference maps a set of symbolic
tensor variables to one raveled vector and back so that a joint log-density over several named parameters can be
sampled as a single position vector.  There are no symbolic variables here, so a parameter is described by a
template (anything with ``shape`` and ``dtype``: NumPy array, torch tensor) or a :class:`ParamSpec`; the map
is batched over chains: a leading ``chains`` axis on every parameter maps to the leading axis of ``q``.
Pure host-side indexing -- no kernel is involved; torch tensors stay on their device.
"""
from __future__ import annotations

from typing import Dict, Iterable, NamedTuple, Sequence, Tuple

import numpy as np

try:  # torch is optional for this module
    import torch
except Exception:  # pragma: no cover
    torch = None


class ParamSpec(NamedTuple):
    name: str
    shape: Tuple[int, ...]
    dtype: object = np.float64


def _np_dtype(dtype):
    if torch is not None and isinstance(dtype, torch.dtype):
        return np.dtype(str(dtype).replace("torch.", ""))
    return np.dtype(dtype)


def _spec(p, index):
    if isinstance(p, ParamSpec):
        return p
    name = getattr(p, "name", None) or f"p{index}"
    shape = tuple(int(s) for s in getattr(p, "shape", ()))
    dtype = getattr(p, "dtype", np.float64)
    return ParamSpec(str(name), shape, dtype)


class RaveledParamsMap:
    """Bidirectional map between a list of named, shaped parameters and one flat vector per chain (the job of the
    reference class of the same name, utils.py:22-74).

    ``ref_params``: iterable of templates or ``ParamSpec``s; the objects themselves are the keys of the dict
    returned by :meth:`unravel_params` (like the reference, which keys by the reference variables)."""

    def __init__(self, ref_params: Iterable):
        self.ref_params = tuple(ref_params)
        self.specs = tuple(_spec(p, i) for i, p in enumerate(self.ref_params))
        # dict keys of unravel_params: the reference objects themselves, or their names when they are unhashable
        # templates (NumPy arrays)
        self.keys = tuple(p if getattr(p, "__hash__", None) else s.name for p, s in zip(self.ref_params, self.specs))
        self.ref_shapes = [s.shape for s in self.specs]
        self.ref_dtypes = [s.dtype for s in self.specs]
        sizes = [int(np.prod(s, dtype=np.int64)) for s in self.ref_shapes]
        ends = np.cumsum(sizes).tolist()
        self.slice_indices = list(zip([0] + ends[:-1], ends))
        self.vec_slices = [slice(*idx) for idx in self.slice_indices]
        self.size = ends[-1] if ends else 0

    def ravel_params(self, params: Sequence):
        """Flatten every parameter and join the pieces in declaration order (reference utils.py:54-56).  Every parameter has
        either its reference shape (result ``[dim]``) or one extra leading chains axis (result ``[chains, dim]``)."""
        if len(params) != len(self.specs):
            raise ValueError(f"expected {len(self.specs)} parameters, got {len(params)}")
        is_torch = torch is not None and any(isinstance(p, torch.Tensor) for p in params)
        flat, batch = [], None
        for p, spec in zip(params, self.specs):
            a = p if (is_torch and isinstance(p, torch.Tensor)) else np.asarray(p)
            nd = len(spec.shape)
            if tuple(a.shape) == spec.shape:
                flat.append((a.reshape(-1), False))
            elif a.ndim == nd + 1 and tuple(a.shape[1:]) == spec.shape:
                if batch is not None and batch != a.shape[0]:
                    raise ValueError("parameters disagree on the number of chains")
                batch = int(a.shape[0])
                flat.append((a.reshape(batch, -1), True))
            else:
                raise ValueError(f"parameter {spec.name}: shape {tuple(a.shape)} does not match {spec.shape}")
        if is_torch:
            dev = next(p.device for p in params if isinstance(p, torch.Tensor))
            cols = []
            for a, batched in flat:
                t = a if isinstance(a, torch.Tensor) else torch.as_tensor(a, device=dev)
                t = t.to(device=dev, dtype=torch.float64)
                cols.append(t if batched or batch is None else t.unsqueeze(0).expand(batch, -1))
            return torch.cat(cols, dim=-1)
        cols = []
        for a, batched in flat:
            a = a.astype(np.float64)
            cols.append(a if batched or batch is None else np.broadcast_to(a, (batch, a.shape[0])))
        return np.concatenate(cols, axis=-1)

    def unravel_params(self, raveled_params) -> Dict[object, object]:
        """Cut a flat vector back into the declared parameters (reference utils.py:58-71): ``[dim]`` gives the
        reference shapes, ``[chains, dim]`` gives ``[chains, *shape]``; values are cast to the reference dtypes."""
        q = raveled_params
        is_torch = torch is not None and isinstance(q, torch.Tensor)
        if not is_torch:
            q = np.asarray(q)
        if q.shape[-1] != self.size:
            raise ValueError(f"raveled vector has {q.shape[-1]} entries, the map needs {self.size}")
        lead = tuple(q.shape[:-1])
        out = {}
        for key, slc, shape, dtype in zip(self.keys, self.vec_slices, self.ref_shapes, self.ref_dtypes):
            v = q[..., slc].reshape(lead + tuple(shape))
            if is_torch:
                out[key] = v.to(dtype if isinstance(dtype, torch.dtype) else getattr(torch, _np_dtype(dtype).name))
            else:
                out[key] = v.astype(_np_dtype(dtype))
        return out

    def __repr__(self):
        return f"{type(self).__name__}(({', '.join(s.name for s in self.specs)}))"
