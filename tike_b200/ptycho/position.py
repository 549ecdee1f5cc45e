"""Scan-position options, bounds check and the per-epoch affine regularisation
(reference: src/tike/ptycho/position.py:137-377, 491-628, 715-776).

The per-position gradient sums are produced by the fused lstsq kernel
(csrc/rpie.cu, reference lstsq.py:545-579); everything here is O(positions)
host arithmetic on (P, 2) arrays.
"""
from __future__ import annotations

import dataclasses
import logging
import typing

import numpy as np

from .. import precision
from .. import random as tb_random
from .._array import to_device, to_host

logger = logging.getLogger(__name__)


@dataclasses.dataclass(frozen=True)
class AffineTransform:
    """2-D affine map: scale @ shear @ rotate, then translate
    (position.py:137-252)."""

    scale0: float = 1.0
    scale1: float = 1.0
    shear1: float = 0.0
    angle: float = 0.0
    t0: float = 0.0
    t1: float = 0.0

    def resample(self, factor: float) -> "AffineTransform":
        return AffineTransform(self.scale0, self.scale1, self.shear1,
                               self.angle, self.t0 * factor, self.t1 * factor)

    @classmethod
    def frombuffer(cls, buffer) -> "AffineTransform":
        return AffineTransform(*buffer)

    def astuple(self) -> tuple:
        return (self.scale0, self.scale1, self.shear1, self.angle, self.t0,
                self.t1)

    def asbuffer(self) -> np.ndarray:
        return np.array(self.astuple())

    @classmethod
    def fromarray(cls, T) -> "AffineTransform":
        """Decompose a 2x2 (or 3x2 with translation row) matrix
        (Graphics Gems II §7.1, position.py:166-192)."""
        T = np.asarray(T)
        R = T[:2, :2].copy()
        scale0 = np.linalg.norm(R[0])
        if scale0 <= 0:
            return AffineTransform()
        R[0] /= scale0
        shear1 = R[0] @ R[1]
        R[1] -= shear1 * R[0]
        scale1 = np.linalg.norm(R[1])
        if scale1 <= 0:
            return AffineTransform()
        R[1] /= scale1
        shear1 /= scale1
        angle = np.arccos(R[0, 0])
        has_t = T.shape[0] > 2
        return AffineTransform(float(scale0), float(scale1), float(shear1),
                               float(angle),
                               float(T[2, 0]) if has_t else 0.0,
                               float(T[2, 1]) if has_t else 0.0)

    def asarray(self, xp=np) -> np.ndarray:
        c, s = np.cos(self.angle), np.sin(self.angle)
        f = precision.floating
        scale = np.array([[self.scale0, 0.0], [0.0, self.scale1]], dtype=f)
        shear = np.array([[1.0, 0.0], [self.shear1, 1.0]], dtype=f)
        rot = np.array([[c, -s], [s, c]], dtype=f)
        return scale @ shear @ rot

    def asarray3(self, xp=np) -> np.ndarray:
        T = np.empty((3, 2), dtype=precision.floating)
        T[2] = (self.t0, self.t1)
        T[:2, :2] = self.asarray()
        return T

    def __call__(self, x, gpu=False, shift=True):
        result = x @ self.asarray()
        if shift:
            result = result + np.array((self.t0, self.t1))
        return result


def _lstsq(a, b, weights=None):
    """inv(a^H a) a^H b (linalg.py:33-64)."""
    if weights is not None:
        a = a * np.sqrt(weights[..., None])
        b = b * np.sqrt(weights[..., None])
    aT = a.conj().swapaxes(-1, -2)
    return np.linalg.inv(aT @ a) @ aT @ b


def estimate_global_transformation(positions0, positions1, weights,
                                   transform=None):
    """Weighted least-squares affine fit (position.py:255-274)."""
    try:
        result = AffineTransform.fromarray(
            _lstsq(a=np.pad(positions0, ((0, 0), (0, 1)), constant_values=1),
                   b=positions1, weights=weights))
    except np.linalg.LinAlgError:
        result = AffineTransform()
    # Frobenius norm without np.linalg.norm: for a 2-D input that goes through
    # a BLAS dot, which is milliseconds per call on threaded BLAS builds
    residual = result(positions0) - positions1
    return result, float(np.sqrt(np.sum(np.square(residual), dtype=np.float64)))


def _native_inliers():
    """tb_affine_inliers of libtikeb200 as a Python callable, or None when the
    library has not been built (host-side bookkeeping also works without it)."""
    import ctypes
    from .. import _lib
    try:
        fn = _lib.lib().tb_affine_inliers
    except _lib.LibraryNotBuilt:
        return None

    def run(x0, y0, x1, y1, t, max_error, mask):
        m = np.ascontiguousarray(t.asarray().astype(np.float64))
        count = ctypes.c_int64(0)
        _lib.check(fn(x0.ctypes.data, y0.ctypes.data, x1.ctypes.data, y1.ctypes.data,
                      len(x0), m.ctypes.data, float(t.t0), float(t.t1),
                      float(max_error) * float(max_error), mask.ctypes.data,
                      ctypes.byref(count)), 'affine inliers')
        return int(count.value)

    return run


def estimate_global_transformation_ransac(positions0, positions1, weights=None,
                                          transform=AffineTransform(),
                                          min_sample: int = 4,
                                          max_error: float = 32,
                                          min_consensus: float = 0.75,
                                          max_iter: int = 20):
    """RANSAC affine fit (position.py:277-327).  Draws its subsets from
    tike_b200.random.randomizer_np like the reference does from
    tike.random.randomizer_np, so seeded runs consume the generator
    identically."""
    best_fitness = np.inf
    subsets = tb_random.randomizer_np.choice(
        a=len(positions0), size=(max_iter, min_sample), replace=True)
    if weights is not None:
        for subset in subsets:
            candidate, _ = estimate_global_transformation(
                positions0[subset], positions1[subset], weights, transform)
            error = np.linalg.norm(candidate(positions0) - positions1, axis=-1)
            inliers = error <= max_error
            if np.sum(inliers) / len(inliers) >= min_consensus:
                candidate, fitness = estimate_global_transformation(
                    positions0[inliers], positions1[inliers], weights, candidate)
                if fitness < best_fitness:
                    best_fitness = fitness
                    transform = candidate
        return transform, best_fitness
    # Unweighted case (what the solvers use): same arithmetic on contiguous
    # coordinate vectors -- (n, 2) arrays make every NumPy reduction walk rows
    # of two elements, which dominated the epoch at 1e5 positions.
    x0 = np.ascontiguousarray(positions0[:, 0], dtype=np.float64)
    y0 = np.ascontiguousarray(positions0[:, 1], dtype=np.float64)
    x1 = np.ascontiguousarray(positions1[:, 0], dtype=np.float64)
    y1 = np.ascontiguousarray(positions1[:, 1], dtype=np.float64)

    def residuals(t: AffineTransform):
        m = t.asarray().astype(np.float64)
        return (x0 * m[0, 0] + y0 * m[1, 0] + t.t0 - x1,
                x0 * m[0, 1] + y0 * m[1, 1] + t.t1 - y1)

    # The per-iteration pass over all positions runs in the native library
    # when it is built (tb_affine_inliers: the same float64 expressions as
    # residuals() below, bit-exact, one pass instead of a dozen NumPy ones).
    native = _native_inliers()
    mask = np.empty(len(x0), dtype=np.uint8)
    everyone = None  # (candidate, fitness) of the fit over ALL points, computed once
    for subset in subsets:
        candidate, _ = estimate_global_transformation(
            positions0[subset], positions1[subset], None, transform)
        if native is not None:
            count = native(x0, y0, x1, y1, candidate, max_error, mask)
            inliers = mask.view(np.bool_)
        else:
            rx, ry = residuals(candidate)
            inliers = (rx * rx + ry * ry) <= max_error * max_error
            count = int(np.count_nonzero(inliers))
        if count == len(inliers):
            # the usual case (32 px is a generous threshold): every iteration
            # refits the same point set, so the refit is done once
            if everyone is None:
                try:
                    fit = AffineTransform.fromarray(
                        _lstsq(a=np.pad(positions0, ((0, 0), (0, 1)), constant_values=1),
                               b=positions1))
                except np.linalg.LinAlgError:
                    fit = AffineTransform()
                rx, ry = residuals(fit)
                everyone = (fit, float(np.sqrt(np.sum(rx * rx + ry * ry))))
            candidate, fitness = everyone
            if fitness < best_fitness:
                best_fitness = fitness
                transform = candidate
            continue
        if count / len(inliers) >= min_consensus:
            # the fit itself keeps the reference's float32 normal equations
            # (linalg.py:33-64): AffineTransform.fromarray recovers the angle
            # with arccos near 1, so the rounding of the fit is visible in the
            # regularised trajectory
            try:
                candidate = AffineTransform.fromarray(
                    _lstsq(a=np.pad(positions0[inliers], ((0, 0), (0, 1)), constant_values=1),
                           b=positions1[inliers]))
            except np.linalg.LinAlgError:
                candidate = AffineTransform()
            rx, ry = residuals(candidate)
            fitness = float(np.sqrt(np.sum((rx * rx + ry * ry)[inliers])))
            if fitness < best_fitness:
                best_fitness = fitness
                transform = candidate
    return transform, best_fitness


@dataclasses.dataclass
class PositionOptions:
    """Settings and state of position correction (position.py:330-588)."""

    initial_scan: np.ndarray
    use_adaptive_moment: bool = False
    vdecay: float = 0.999
    mdecay: float = 0.9
    use_position_regularization: bool = False
    update_magnitude_limit: float = 0
    transform: AffineTransform = AffineTransform()
    origin: typing.Any = dataclasses.field(
        init=True, default_factory=lambda: np.zeros(2))
    confidence: typing.Any = dataclasses.field(
        init=True, default_factory=lambda: None)
    update_start: int = 0
    _momentum: typing.Any = dataclasses.field(
        init=False, default_factory=lambda: None)

    def __post_init__(self):
        host = to_host(self.initial_scan)
        if isinstance(self.initial_scan, np.ndarray):
            self.initial_scan = host.astype(precision.floating)
        if self.confidence is None:
            self.confidence = np.ones(shape=host.shape,
                                      dtype=precision.floating)
        if self.use_adaptive_moment:
            self._momentum = np.zeros((*host.shape[:-1], 4),
                                      dtype=precision.floating)

    def _like(self, initial_scan, **kw) -> "PositionOptions":
        base = dict(
            use_adaptive_moment=self.use_adaptive_moment, vdecay=self.vdecay,
            mdecay=self.mdecay,
            use_position_regularization=self.use_position_regularization,
            update_magnitude_limit=self.update_magnitude_limit,
            transform=self.transform, update_start=self.update_start,
            origin=self.origin)
        base.update(kw)
        return PositionOptions(initial_scan, **base)

    def split(self, indices) -> "PositionOptions":
        new = self._like(self.initial_scan[..., indices, :])
        if self.confidence is not None:
            new.confidence = self.confidence[..., indices, :]
        if self.use_adaptive_moment:
            new._momentum = self._momentum[..., indices, :]
        return new

    @staticmethod
    def join(x, reorder):
        if any(e is None for e in x):
            return None
        new = x[0]._like(
            np.concatenate([to_host(e.initial_scan) for e in x], axis=0)[reorder])
        if x[0].confidence is not None:
            new.confidence = np.concatenate(
                [to_host(e.confidence) for e in x], axis=0)[reorder]
        if x[0].use_adaptive_moment:
            new._momentum = np.concatenate(
                [to_host(e._momentum) for e in x], axis=0)[reorder]
        return new

    def copy_to_device(self) -> "PositionOptions":
        # (P, 2) bookkeeping stays on the host: the position step is a
        # per-epoch O(P) operation (see solvers/lstsq.py)
        new = self._like(to_host(self.initial_scan), confidence=to_host(self.confidence))
        if self.use_adaptive_moment:
            new._momentum = to_host(self._momentum).astype(precision.floating)
        return new

    def copy_to_host(self) -> "PositionOptions":
        new = self._like(to_host(self.initial_scan), confidence=to_host(self.confidence))
        if self.use_adaptive_moment:
            new._momentum = to_host(self._momentum)
        return new

    def resample(self, factor: float) -> "PositionOptions":
        return self._like(self.initial_scan * factor,
                          transform=self.transform.resample(factor),
                          confidence=self.confidence,
                          origin=np.asarray(self.origin) * factor)

    @property
    def v(self):
        return self._momentum[..., 0:2]

    @v.setter
    def v(self, x):
        self._momentum[..., 0:2] = x

    @property
    def m(self):
        return self._momentum[..., 2:4]

    @m.setter
    def m(self, x):
        self._momentum[..., 2:4] = x


def check_allowed_positions(scan, psi, probe_shape: tuple):
    """Positions must leave a one-pixel margin inside psi
    (position.py:600-628)."""
    scan = to_host(scan)
    int_scan = scan // 1
    lo = np.min(int_scan, axis=-2)
    hi = np.max(int_scan, axis=-2)
    hi_ok = (psi.shape[-2] - probe_shape[-2] - 1,
             psi.shape[-1] - probe_shape[-1] - 1)
    if lo[0] < 1 or lo[1] < 1 or hi[0] > hi_ok[0] or hi[1] > hi_ok[1]:
        raise ValueError(
            "Scan positions must be >= 1 and "
            "scan positions + 1 + probe.shape must be <= psi.shape. "
            "psi may be too small or the scan positions may be scaled wrong. "
            f"The span of scan is {lo} to {hi}, and "
            f"the shape of psi is {tuple(psi.shape)}.")


def affine_position_regularization(updated, position_options: PositionOptions,
                                   max_error: float = 32):
    """Fit a global affine transform to the position updates every epoch and
    optionally pull positions toward it (position.py:731-776)."""
    host = to_host(updated)
    p0 = to_host(position_options.initial_scan)
    origin = np.asarray(to_host(position_options.origin))
    new_transform, _ = estimate_global_transformation_ransac(
        positions0=p0 - origin, positions1=host - origin,
        transform=position_options.transform, max_error=max_error)
    position_options.transform = new_transform
    if position_options.use_position_regularization:
        relax = 0.9
        predicted = position_options.transform(p0, shift=False)
        host = (host * (1 - relax) + relax * predicted).astype(host.dtype)
        if not isinstance(updated, np.ndarray):
            return to_device(host, device=updated.device), position_options
        return host, position_options
    return updated, position_options


def gaussian_gradient_taps(sigma: float = 0.333) -> np.ndarray:
    """Correlation taps t[-2..2] of scipy.ndimage.gaussian_filter1d(order=1,
    truncate=6) used by the reference's gaussian_gradient
    (position.py:779-810): derivative(i) = sum_t taps[t+2] * x[i+t]."""
    import scipy.ndimage
    radius = int(6.0 * sigma + 0.5)
    if radius != 2:
        raise ValueError('the fused kernel carries 5 taps (sigma = 0.333)')
    impulse = np.zeros(2 * radius + 1, dtype=np.float64)
    impulse[radius] = 1
    response = scipy.ndimage.gaussian_filter1d(
        impulse, sigma=sigma, order=1, mode='constant', truncate=6.0)
    return response[::-1].astype(np.float32)


def gaussian_gradient(x, sigma: float = 0.333):
    """First-order Gaussian derivatives of the last two axes of ``x`` with the
    sign convention of the reference (position.py:779-810): the filter is
    applied to ``-x``, edges replicate.  NumPy arrays and torch tensors."""
    import scipy.ndimage
    if isinstance(x, np.ndarray):
        return tuple(
            scipy.ndimage.gaussian_filter1d(-x, sigma=sigma, order=1, axis=axis,
                                            mode='nearest', truncate=6.0)
            for axis in (-2, -1))
    import torch
    radius = int(6.0 * sigma + 0.5)
    impulse = np.zeros(2 * radius + 1)
    impulse[radius] = 1
    taps = scipy.ndimage.gaussian_filter1d(impulse, sigma=sigma, order=1,
                                           mode='constant', truncate=6.0)[::-1].copy()
    w = torch.as_tensor(taps, dtype=torch.float32, device=x.device)
    out = []
    for axis in (-2, -1):
        n = x.shape[axis]
        index = torch.arange(n, device=x.device)
        acc = torch.zeros_like(x)
        for t in range(-radius, radius + 1):
            acc = acc - w[t + radius] * x.index_select(axis % x.ndim,
                                                       (index + t).clamp(0, n - 1))
        out.append(acc)
    return tuple(out)
