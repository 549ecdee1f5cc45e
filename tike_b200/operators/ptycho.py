"""Ptychography forward operator (reference:
src/tike/operators/cupy/ptycho.py:26-204, multislice.py:18-279).

``Ptycho = Propagation o Multislice(Convolution(Patch))``.  The fused C-ABI
path is used whenever the default Propagation / Multislice classes are
composed (single-slice objects: one kernel; D > 1 slices: the chunked
multislice driver, csrc/multislice.cu), otherwise the generic operator
composition runs.
"""
from __future__ import annotations

import typing

import torch

from .. import kernels
from .._array import to_device
from . import objective
from .convolution import Convolution
from .operator import Operator
from .propagation import FresnelSpectProp, Propagation


class Multislice(Operator):
    """Object-probe interaction through D slices with a Fresnel-spectrum
    step between them (multislice.py:18-194).  With one slice the inter-slice
    propagator is never applied."""

    def __init__(self, detector_shape: int, probe_shape: int,
                 probe_wavelength: float = float('nan'),
                 probe_FOV_lengths=(float('nan'), float('nan')), nz: int = 0,
                 n: int = 0, multislice_propagation_distance: float = 0.0,
                 propagation=FresnelSpectProp, diffraction=Convolution,
                 norm: str = "ortho", **kwargs):
        self.diffraction = diffraction(probe_shape=probe_shape,
                                       detector_shape=detector_shape, nz=nz,
                                       n=n, **kwargs)
        self.propagation = (propagation or FresnelSpectProp)(
            norm=norm, probe_shape=probe_shape, wavelength=probe_wavelength,
            probe_FOV=probe_FOV_lengths,
            distance=multislice_propagation_distance, **kwargs)
        self.probe_shape = probe_shape
        self.detector_shape = detector_shape
        self.nz, self.n = nz, n
        self.probe_wavelength = probe_wavelength
        self.probe_FOV_lengths = probe_FOV_lengths
        self.multislice_propagation_distance = multislice_propagation_distance

    def __enter__(self):
        self.diffraction.__enter__()
        return self

    def __exit__(self, type, value, traceback):
        self.diffraction.__exit__(type, value, traceback)

    @staticmethod
    def _check_psi(psi):
        if psi.ndim != 3:
            raise ValueError(f'psi must have shape (D, H, W), not {tuple(psi.shape)}')

    def fwd(self, probe, scan, psi, **kwargs):
        psi = to_device(psi, dtype='c64')
        self._check_psi(psi)
        exitwave = self.diffraction.fwd(psi=psi[0], scan=scan, probe=probe)
        for s in range(1, psi.shape[0]):
            exitwave = self.diffraction.fwd(
                psi=psi[s], scan=scan, probe=self.propagation.fwd(exitwave))
        return exitwave

    def fwd_return_intermediate_probes(self, probe, scan, psi, **kwargs):
        psi = to_device(psi, dtype='c64')
        probe = to_device(probe, dtype='c64')
        self._check_psi(psi)
        p = probe[..., 0, :, :, :] if probe.ndim == 5 else probe
        incident = p.expand(scan.shape[-2], *p.shape[-3:])
        probes = [incident]
        for t in range(psi.shape[0]):
            exitwave = self.diffraction.fwd(psi=psi[t], scan=scan, probe=probes[t])
            if t == psi.shape[0] - 1:
                break
            probes.append(self.propagation.fwd(nearplane=exitwave))
        return exitwave, torch.stack([q.expand_as(probes[-1]) for q in probes])

    def adj(self, nearplane, probe, scan, psi, overwrite=False, **kwargs):
        psi = to_device(psi, dtype='c64')
        self._check_psi(psi)
        nslices = int(psi.shape[0])
        # probe incident on every slice (multislice.py:151-163)
        probes = [probe]
        for t in range(1, nslices):
            probes.append(self.propagation.fwd(
                self.diffraction.fwd(psi=psi[t - 1], scan=scan, probe=probes[t - 1])))
        # back through the slices: object adjoint of slice t, then the wave that
        # left slice t - 1 (multislice.py:164-192).  The object map is
        # homogeneous of degree `nslices`, hence the division (Euler) that
        # makes <fwd(psi), y> == <psi, adj(y)>.
        psi_adj = [None] * nslices
        wave = nearplane
        for t in range(nslices - 1, -1, -1):
            psi_adj[t] = self.diffraction.adj(nearplane=wave, probe=probes[t], scan=scan,
                                              overwrite=False)
            wave = self.diffraction.adj_probe(nearplane=wave, scan=scan, psi=psi[t])
            if t > 0:
                wave = self.propagation.adj(wave)
        return torch.stack(psi_adj) / nslices, wave

    @property
    def patch(self):
        return self.diffraction.patch

    @property
    def pad(self):
        return self.diffraction.pad

    @property
    def end(self):
        return self.diffraction.end


SingleSlice = Multislice


class Ptycho(Operator):
    """farplane = Propagation(probe * patches(psi, scan)).

    probe (1|POSI, 1, SHARED, W, H); psi (1, WIDE, HIGH); scan (POSI, 2);
    farplane (POSI, 1, SHARED, detector, detector).
    """

    def __init__(self, detector_shape: int, probe_shape: int,
                 probe_wavelength: float = float('nan'),
                 probe_FOV_lengths=(float('nan'), float('nan')), nz: int = 0,
                 n: int = 0, multislice_propagation_distance: float = 1e-9,
                 propagation: typing.Type[Propagation] = Propagation,
                 diffraction: typing.Type[Multislice] = Multislice,
                 norm: str = 'ortho', **kwargs):
        self.propagation = propagation(detector_shape=detector_shape,
                                       norm=norm, **kwargs)
        self.diffraction = diffraction(
            probe_shape=probe_shape, probe_wavelength=probe_wavelength,
            probe_FOV_lengths=probe_FOV_lengths, detector_shape=detector_shape,
            nz=nz, n=n,
            multislice_propagation_distance=multislice_propagation_distance,
            **kwargs)
        self.probe_shape = probe_shape
        self.detector_shape = detector_shape
        self.nz, self.n = nz, n
        self.norm = norm
        self.probe_wavelength = probe_wavelength
        self.probe_FOV_lengths = probe_FOV_lengths
        self.multislice_propagation_distance = multislice_propagation_distance
        self._fused = (type(self.propagation) is Propagation
                       and type(self.diffraction) is Multislice
                       and type(self.diffraction.diffraction) is Convolution)

    def __enter__(self):
        self.propagation.__enter__()
        self.diffraction.__enter__()
        return self

    def __exit__(self, type, value, traceback):
        self.propagation.__exit__(type, value, traceback)
        self.diffraction.__exit__(type, value, traceback)

    def _fused_fwd(self, probe, scan, psi, want_farplane=True,
                   want_intensity=False):
        psi = to_device(psi, dtype='c64')
        scan = to_device(scan, dtype='f32')
        probe = to_device(probe, dtype='c64')
        Multislice._check_psi(psi)
        if probe.ndim != 5 or probe.shape[1] != 1:
            raise ValueError(f'probe must be (1|POSI, 1, S, W, H), not {tuple(probe.shape)}')
        p = probe[:, 0]
        if p.shape[0] == 1:
            p = p[0]
        B, M, D = scan.shape[0], probe.shape[-3], self.detector_shape
        nslices = int(psi.shape[0])
        far = torch.empty((B, 1, M, D, D), dtype=torch.complex64,
                          device=psi.device) if want_farplane or (D not in (16, 32, 64, 128) and nslices == 1) else None
        inten = torch.empty((B, D, D), dtype=torch.float32,
                            device=psi.device) if want_intensity else None
        if nslices == 1:
            batch = kernels.make_batch(psi[0], scan, p.contiguous(), D, self.norm)
            kernels.ptycho_fwd(batch, far, inten)
        else:
            batch = kernels.multislice_batch(psi.contiguous(), scan, p.contiguous(), D, self.norm)
            kernels.multislice_fwd(batch, nslices, self.fresnel_propagator(psi.device),
                                   far, inten, device=psi.device)
        return far, inten

    def fresnel_propagator(self, device):
        """Inter-slice Fresnel kernel (probe_shape, probe_shape) on ``device``."""
        prop = self.diffraction.propagation
        if not isinstance(prop, FresnelSpectProp):
            raise NotImplementedError(
                'the fused multislice path needs the FresnelSpectProp inter-slice operator')
        return prop.propagator(device, self.probe_shape)

    def fwd(self, probe, scan, psi, **kwargs):
        if self._fused:
            return self._fused_fwd(probe, scan, psi)[0]
        probe = to_device(probe, dtype='c64')
        return self.propagation.fwd(
            self.diffraction.fwd(psi=psi, scan=scan,
                                 probe=probe[..., 0, :, :, :]),
            overwrite=True)[..., None, :, :, :]

    def fwd_return_intermediate_probes(self, probe, scan, psi, **kwargs):
        probe = to_device(probe, dtype='c64')
        far = self.fwd(probe=probe, scan=scan, psi=psi)
        p = probe[..., 0, :, :, :]
        return far, p.expand(scan.shape[-2], *p.shape[-3:])[None]

    def adj(self, farplane, probe, scan, psi, overwrite=False, **kwargs):
        probe = to_device(probe, dtype='c64')
        near = self.propagation.adj(farplane, overwrite=overwrite)[..., 0, :, :, :]
        psi_adj, probe_adj = self.diffraction.adj(
            nearplane=near, probe=probe[..., 0, :, :, :], scan=scan,
            overwrite=True, psi=psi)
        return psi_adj, probe_adj[..., None, :, :, :]

    def _compute_intensity(self, data, psi, scan, probe):
        if self._fused:
            far, inten = self._fused_fwd(probe, scan, psi, want_farplane=True,
                                         want_intensity=True)
            return inten, far
        far = self.fwd(psi=psi, scan=scan, probe=probe)
        return torch.sum((far * far.conj()).real,
                         dim=tuple(range(1, far.ndim - 2))), far

    def intensity(self, psi, scan, probe):
        """Detector intensity only; the far-field wave stays on chip."""
        if self._fused and self.detector_shape in (16, 32, 64, 128):
            return self._fused_fwd(probe, scan, psi, want_farplane=False,
                                   want_intensity=True)[1]
        return self._compute_intensity(None, psi, scan, probe)[0]

    def cost(self, data, psi, scan, probe, *, model: str):
        intensity = self.intensity(psi, scan, probe)
        return getattr(objective, model)(to_device(data, dtype='f32'), intensity)
