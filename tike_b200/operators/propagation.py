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


class FresnelSpectProp(Propagation):
    """Inter-slice propagation by the Fresnel spectrum method
    (reference: src/tike/operators/cupy/fresnelspectprop.py:17-135):
    ``fwd = ifft2(fft2(x) * H)``, ``adj = ifft2(fft2(x) * conj(H))`` with
    ``H = fftshift(exp(i z sqrt(k^2 - kx^2 - ky^2)))``."""

    def __init__(self, norm: str = "ortho", probe_shape: int = 0,
                 wavelength: float = 1e-9, probe_FOV=(1e-6, 1e-6),
                 distance: float = 1e-6, detector_shape: int = 0, **kwargs):
        self.norm = norm
        self.detector_shape = probe_shape or detector_shape
        self.probe_FOV = probe_FOV
        self.distance = distance
        self.wavelength = wavelength
        self._cache = {}

    def propagator(self, device, n: int = 0):
        """(n, n) complex64 kernel on ``device`` (cached)."""
        n = int(n or self.detector_shape)
        key = (n, str(device))
        if key not in self._cache:
            import torch
            host = kernels.fresnel_propagator(n, self.probe_FOV, self.distance,
                                              self.wavelength)
            self._cache[key] = torch.from_numpy(host).to(device)
        return self._cache[key]

    def _apply(self, x, conj: bool, overwrite: bool):
        x = to_device(x, dtype='c64')
        if x.shape[-1] != x.shape[-2]:
            raise ValueError(f"waves must be square, not {tuple(x.shape)}.")
        if not overwrite:
            x = x.clone()
        n = int(x.shape[-1])
        h = self.propagator(x.device, n)
        x = kernels.fft2(x.contiguous(), inverse=False, scale=1.0)
        x *= (h.conj() if conj else h)
        return kernels.fft2(x, inverse=True, scale=1.0 / (n * n))

    def fwd(self, nearplane, overwrite: bool = False, **kwargs):
        return self._apply(nearplane, False, overwrite)

    def adj(self, farplane, overwrite: bool = False, **kwargs):
        return self._apply(farplane, True, overwrite)
