"""Object (psi) options and per-epoch constraints
(reference: src/tike/ptycho/object.py)."""
from __future__ import annotations

import copy
import dataclasses
import logging
import typing

import numpy as np
import torch

from .. import linalg, precision
from .._array import to_device, to_host

logger = logging.getLogger(__name__)


@dataclasses.dataclass
class ObjectOptions:
    """Settings and state of the object update (object.py:25-81)."""

    convergence_tolerance: float = 0
    update_mnorm: typing.List[float] = dataclasses.field(
        init=False, default_factory=list)
    positivity_constraint: float = 0
    smoothness_constraint: float = 0
    use_adaptive_moment: bool = False
    vdecay: float = 0.999
    mdecay: float = 0.9
    v: typing.Any = dataclasses.field(init=False, default=None)
    m: typing.Any = dataclasses.field(init=False, default=None)
    preconditioner: typing.Any = dataclasses.field(init=False, default=None)
    clip_magnitude: bool = False
    multislice_propagation_distance: float = 1.0e-9

    def _clone(self) -> "ObjectOptions":
        o = ObjectOptions(
            convergence_tolerance=self.convergence_tolerance,
            positivity_constraint=self.positivity_constraint,
            smoothness_constraint=self.smoothness_constraint,
            use_adaptive_moment=self.use_adaptive_moment,
            vdecay=self.vdecay,
            mdecay=self.mdecay,
            clip_magnitude=self.clip_magnitude,
            multislice_propagation_distance=self.multislice_propagation_distance,
        )
        o.update_mnorm = copy.copy(self.update_mnorm)
        return o

    def copy_to_device(self) -> "ObjectOptions":
        o = self._clone()
        o.v, o.m = to_device(self.v), to_device(self.m)
        o.preconditioner = to_device(self.preconditioner, dtype='c64')
        return o

    def copy_to_host(self) -> "ObjectOptions":
        o = self._clone()
        o.v, o.m = to_host(self.v), to_host(self.m)
        o.preconditioner = to_host(self.preconditioner)
        return o

    def resample(self, factor: float, interp) -> "ObjectOptions":
        return self._clone()  # momentum restarts when the grid changes

    @staticmethod
    def join_psi(x, stripe_start, probe_width: int):
        """Stitch per-worker objects by stripes (object.py:154-167)."""
        joined = x[0]
        w = probe_width // 2
        for i in range(1, len(x)):
            lo = stripe_start[i] + w
            hi = stripe_start[i + 1] + w if i + 1 < len(x) else x[0].shape[1]
            joined[:, lo:hi, :] = x[i][:, lo:hi, :]
        return joined

    @staticmethod
    def join(x, stripe_start, probe_width: int) -> "ObjectOptions":
        o = x[0]._clone()
        for name in ('v', 'm', 'preconditioner'):
            if getattr(x[0], name) is not None:
                setattr(o, name, ObjectOptions.join_psi(
                    [getattr(e, name) for e in x], stripe_start, probe_width))
        return o


def positivity_constraint(x, r: float):
    """r * |x| + (1 - r) * x (object.py:208-224)."""
    if r > 0:
        if r > 1:
            raise ValueError(
                f"Positivity constraint must be in the range [0, 1] not {r}.")
        return (r * x.abs() + (1 - r) * x).to(x.dtype)
    return x


def smoothness_constraint(x, a: float):
    """3x3 box-like smoothing with edge replication (object.py:227-253)."""
    if not (0 <= a < 1.0 / 8.0):
        raise ValueError(
            f"Smoothness constraint must be in range [0, 1/8) not {a}.")
    k = torch.full((3, 3), a, dtype=torch.float32, device=x.device)
    k[1, 1] = 1.0 - 8.0 * a
    lead = x.shape[:-2]

    def conv(t):
        t = t.reshape(-1, 1, *t.shape[-2:])
        t = torch.nn.functional.pad(t, (1, 1, 1, 1), mode='replicate')
        return torch.nn.functional.conv2d(t, k[None, None]).reshape(*lead, *x.shape[-2:])

    return torch.complex(conv(x.real.contiguous()), conv(x.imag.contiguous()))


def clip_magnitude(x, a_max: float = 1.0):
    """Clip |x| to a_max keeping the phase (ptycho.py:257-262)."""
    mag = x.abs()
    scale = torch.where(mag > a_max, a_max / mag, torch.ones_like(mag))
    return x * scale


def remove_object_ambiguity(psi, probe, preconditioner):
    """Fix the psi/probe scale ambiguity (object.py:324-335)."""
    W = preconditioner.real
    W = W / linalg.mnorm(W)
    object_norm = 2 * torch.sqrt(torch.mean(torch.square(psi.abs()) * W))
    return psi / object_norm, probe * object_norm


def get_padded_object(scan, probe, extra: int = 0):
    """0.5-initialised object covering the scan + shifted scan
    (object.py:256-277)."""
    int_scan = scan // 1
    min_corner = np.min(int_scan, axis=-2)
    max_corner = np.max(int_scan, axis=-2)
    span = max_corner - min_corner + probe.shape[-1] + 2 + 2 * extra
    psi = np.full(span.astype(precision.integer), 0.5 + 0j,
                  dtype=precision.cfloating)
    return psi, scan + 1 - min_corner + extra


def get_absorbtion_image(data, scan, *, rescale=1.0, method='cubic'):
    """Scanning-transmission style image from the diffraction patterns: the
    total counts of every pattern interpolated (scipy.interpolate.griddata)
    onto a unit grid spanning the rescaled scan (object.py:281-321); points
    outside the convex hull of the scan take the largest value."""
    import scipy.interpolate
    data = np.asarray(to_host(data), dtype=np.float64)
    points = np.asarray(to_host(scan)) * rescale
    axes = [np.arange(int(np.floor(points[:, d].min())), int(np.ceil(points[:, d].max())))
            for d in (0, 1)]
    rows, cols = np.meshgrid(*axes, indexing='ij')
    counts = np.sum(np.square(data), axis=(-2, -1))
    image = scipy.interpolate.griddata(points=points, values=counts,
                                       xi=(rows.ravel(), cols.ravel()),
                                       method=method, fill_value=np.amax(counts))
    return image.reshape(rows.shape)
