"""Far-field propagation (reference: src/tike/operators/cupy/propagation.py)."""
from __future__ import annotations

from .. import kernels
from .._array import to_device
from .operator import Operator


class Propagation(Operator):
    """2-D FFT of the last two axes, DC at the corner; ``norm`` as in
    scipy.fft ('ortho', 'forward', 'backward')."""

    def __init__(self, detector_shape: int, norm: str = "ortho", **kwargs):
        self.detector_shape = detector_shape
        self.norm = norm

    def _check_shape(self, x) -> None:
        shape = (-1, self.detector_shape, self.detector_shape)
        if tuple(x.shape[-2:]) != shape[-2:]:
            raise ValueError(f"waves must have shape {shape} not {tuple(x.shape)}.")

    def fwd(self, nearplane, overwrite: bool = False, **kwargs):
        x = to_device(nearplane, dtype='c64')
        self._check_shape(x)
        if not overwrite:
            x = x.clone()
        scale, _ = kernels.fft_scales(self.detector_shape, self.norm)
        return kernels.fft2(x, inverse=False, scale=scale)

    def adj(self, farplane, overwrite: bool = False, **kwargs):
        x = to_device(farplane, dtype='c64')
        self._check_shape(x)
        if not overwrite:
            x = x.clone()
        _, scale = kernels.fft_scales(self.detector_shape, self.norm)
        return kernels.fft2(x, inverse=True, scale=scale)


class ZeroPropagation(Propagation):
    """Zero-distance propagation: identity (propagation.py:76-118)."""

    def fwd(self, nearplane, overwrite: bool = False, **kwargs):
        return nearplane

    def adj(self, farplane, overwrite: bool = False, **kwargs):
        return farplane
