"""Patch operator (reference: src/tike/operators/cupy/patch.py:62-188)."""
from __future__ import annotations

import torch

from .. import kernels
from .._array import to_device
from .operator import Operator


def _flat(x, ndim_tail):
    """Collapse leading dims so the kernel sees (nimage, ...)."""
    lead = x.shape[:-ndim_tail]
    n = 1
    for s in lead:
        n *= int(s)
    return x.reshape(n, *x.shape[-ndim_tail:]), lead


class Patch(Operator):
    """Extract (zero-padded) patches from images at sub-pixel positions with
    bilinear interpolation, and add them back (adjoint).

    images (..., H, W) c64; positions (..., N, 2) f32;
    patches (..., N * nrepeat, width+, width+) c64, or (..., L, ...) for the
    adjoint with N * nrepeat = K * L broadcast (patch.py:62-77).
    """

    def fwd(self, images, positions, patches=None, patch_width=0, height=0,
            width=0, nrepeat=1):
        images = to_device(images, dtype='c64')
        positions = to_device(positions, dtype='f32')
        if patches is None and patch_width == 0:
            raise AttributeError('patch_width is required when patches is None')
        patch_width = patches.shape[-1] if patch_width == 0 else patch_width
        if patches is None:
            patches = torch.zeros(
                (*positions.shape[:-2], positions.shape[-2] * nrepeat,
                 patch_width, patch_width), dtype=torch.complex64,
                device=images.device)
        else:
            patches = to_device(patches, dtype='c64')
        assert patch_width <= patches.shape[-1]
        assert images.shape[:-2] == positions.shape[:-2]
        assert positions.shape[:-2] == patches.shape[:-3], (positions.shape,
                                                            patches.shape)
        assert positions.shape[-2] * nrepeat == patches.shape[-3]
        assert positions.shape[-1] == 2, positions.shape
        kernels.patch_fwd(images, positions, patches, patch_width, nrepeat)
        return patches

    def adj(self, positions, patches, images=None, patch_width=0, height=0,
            width=0, nrepeat=1):
        patches = to_device(patches, dtype='c64')
        positions = to_device(positions, dtype='f32')
        patch_width = patches.shape[-1] if patch_width == 0 else patch_width
        assert patch_width <= patches.shape[-1]
        if images is None:
            images = torch.zeros((*positions.shape[:-2], height, width),
                                 dtype=torch.complex64, device=patches.device)
        else:
            images = to_device(images, dtype='c64')
        leading = images.shape[:-2]
        assert positions.shape[:-2] == leading
        N = positions.shape[-2]
        assert positions.shape[-1] == 2
        assert patches.shape[:-3] == leading
        K = patches.shape[-3]
        assert (N * nrepeat) % K == 0 and K >= nrepeat
        assert patches.shape[-1] == patches.shape[-2]
        kernels.patch_adj(images, positions, patches, patch_width, nrepeat)
        return images
