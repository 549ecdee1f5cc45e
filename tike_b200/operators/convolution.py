"""Probe x object-patch product (reference:
src/tike/operators/cupy/convolution.py:11-154)."""
from __future__ import annotations

import torch

from .._array import to_device
from .operator import Operator
from .patch import Patch


class Convolution(Operator):
    """nearplane[s, m] = pad(probe[s|0, m] * patch_s(psi)).

    psi (..., nz, n); probe (..., nscan|1, nprobe, w, w);
    nearplane (..., nscan, nprobe, detector, detector); scan (..., nscan, 2).
    """

    def __init__(self, probe_shape, nz, n, ntheta=None, detector_shape=None,
                 **kwargs):
        self.probe_shape = probe_shape
        self.nz = nz
        self.n = n
        self.detector_shape = probe_shape if detector_shape is None else detector_shape
        self.pad = (self.detector_shape - self.probe_shape) // 2
        self.end = self.probe_shape + self.pad
        self.patch = Patch()

    def fwd(self, psi, scan, probe):
        psi = to_device(psi, dtype='c64')
        scan = to_device(scan, dtype='f32')
        probe = to_device(probe, dtype='c64')
        assert psi.shape[:-2] == scan.shape[:-2], (psi.shape, scan.shape)
        assert probe.shape[:-4] == scan.shape[:-2], (probe.shape, scan.shape)
        assert probe.shape[-4] == 1 or probe.shape[-4] == scan.shape[-2]
        M, D = probe.shape[-3], self.detector_shape
        patches = torch.zeros((*scan.shape[:-2], scan.shape[-2] * M, D, D),
                              dtype=torch.complex64, device=psi.device)
        patches = self.patch.fwd(patches=patches, images=psi, positions=scan,
                                 patch_width=self.probe_shape, nrepeat=M)
        patches = patches.reshape((*scan.shape[:-1], M, D, D))
        patches[..., self.pad:self.end, self.pad:self.end] *= probe
        return patches

    def adj(self, nearplane, scan, probe, psi=None, overwrite=False):
        nearplane = to_device(nearplane, dtype='c64')
        scan = to_device(scan, dtype='f32')
        probe = to_device(probe, dtype='c64')
        assert probe.shape[:-4] == scan.shape[:-2], (probe.shape, scan.shape)
        assert probe.shape[-4] == 1 or probe.shape[-4] == scan.shape[-2]
        assert nearplane.shape[:-3] == scan.shape[:-1], (nearplane.shape, scan.shape)
        if not overwrite:
            nearplane = nearplane.clone()
        nearplane[..., self.pad:self.end, self.pad:self.end] *= probe.conj()
        if psi is None:
            psi = torch.zeros((*scan.shape[:-2], self.nz, self.n),
                              dtype=torch.complex64, device=nearplane.device)
        assert psi.shape[:-2] == scan.shape[:-2]
        return self.patch.adj(
            patches=nearplane.reshape((*scan.shape[:-2],
                                       scan.shape[-2] * nearplane.shape[-3],
                                       *nearplane.shape[-2:])).contiguous(),
            images=psi, positions=scan, patch_width=self.probe_shape,
            nrepeat=nearplane.shape[-3])

    def adj_probe(self, nearplane, scan, psi, overwrite=False):
        nearplane = to_device(nearplane, dtype='c64')
        scan = to_device(scan, dtype='f32')
        psi = to_device(psi, dtype='c64')
        assert nearplane.shape[:-3] == scan.shape[:-1], (nearplane.shape, scan.shape)
        assert psi.shape[:-2] == scan.shape[:-2], (psi.shape, scan.shape)
        M, w = nearplane.shape[-3], self.probe_shape
        patches = torch.zeros((*scan.shape[:-2], scan.shape[-2] * M, w, w),
                              dtype=torch.complex64, device=psi.device)
        patches = self.patch.fwd(patches=patches, images=psi, positions=scan,
                                 patch_width=w, nrepeat=M)
        patches = patches.reshape((*scan.shape[:-1], M, w, w)).conj()
        return patches * nearplane[..., self.pad:self.end, self.pad:self.end]
