"""Operator base class (reference: src/tike/operators/cupy/operator.py:12-57)."""
from __future__ import annotations

import abc

import torch

from .._array import to_device, to_host


class Operator(abc.ABC):
    """Context-manager base of the forward/adjoint operators.  Device arrays
    are torch CUDA tensors (``xp`` is torch); any ``__cuda_array_interface__``
    exporter is accepted as input."""

    xp = torch

    @classmethod
    def asarray(cls, *args, device=None, **kwargs):
        dtype = kwargs.pop('dtype', None)
        t = to_device(args[0], device=None if device is None else
                      torch.device('cuda', device))
        if dtype is not None:
            import numpy as np
            t = t.to({np.dtype('complex64'): torch.complex64,
                      np.dtype('float32'): torch.float32}.get(np.dtype(dtype), t.dtype))
        return t

    @classmethod
    def asnumpy(cls, *args, **kwargs):
        return to_host(args[0])

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        pass

    def fwd(self, **kwargs):
        raise NotImplementedError("The forward operator was not implemented!")

    def adj(self, **kwargs):
        raise NotImplementedError("The adjoint operator was not implemented!")
