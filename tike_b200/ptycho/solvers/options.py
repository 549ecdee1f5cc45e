"""Algorithm options and the parameter container
(reference: src/tike/ptycho/solvers/options.py:19-409, plus DmOptions which
the mounted reference snapshot lacks — SURVEY.md §0 F1)."""
from __future__ import annotations

import abc
import copy
import dataclasses
import typing

import numpy as np
import scipy.ndimage

from ... import precision
from ..._array import to_device, to_host
from ..exitwave import ExitWaveOptions, crop_fourier_space
from ..object import ObjectOptions
from ..position import PositionOptions, check_allowed_positions
from ..probe import ProbeOptions


@dataclasses.dataclass
class IterativeOptions(abc.ABC):
    """Options shared by the iterative solvers (options.py:19-79)."""

    name: str = dataclasses.field(default='', init=False)
    num_batch: int = 1
    batch_method: str = 'wobbly_center'
    rescale_method: str = 'mean_of_abs_object'
    rescale_period: int = 10
    costs: typing.List[typing.List[float]] = dataclasses.field(
        init=False, default_factory=list)
    num_iter: int = 1
    times: typing.List[float] = dataclasses.field(init=False,
                                                  default_factory=list)
    convergence_window: int = 0
    time_limit: float = np.inf


@dataclasses.dataclass
class RpieOptions(IterativeOptions):
    """Regularised ptychographic iterative engine (options.py:82-90)."""

    name: str = dataclasses.field(default='rpie', init=False)
    num_batch: int = 5
    alpha: float = 0.05


@dataclasses.dataclass
class LstsqOptions(IterativeOptions):
    """Least-squares maximum-likelihood solver (options.py:93-95)."""

    name: str = dataclasses.field(default='lstsq_grad', init=False)


@dataclasses.dataclass
class DmOptions(IterativeOptions):
    """Difference-map-style solver: numerators accumulated over all batches,
    one object/probe update per epoch (see solvers/dm.py)."""

    name: str = dataclasses.field(default='dm', init=False)
    num_batch: int = 1


@dataclasses.dataclass
class PtychoParameters:
    """Forward-model parameters (options.py:98-330)."""

    probe: typing.Any
    """(1, 1, SHARED, WIDE, HIGH) complex64 shared illumination."""

    psi: typing.Any
    """(DEPTH, WIDE, HIGH) complex64 object."""

    scan: typing.Any
    """(POSI, 2) float32 minimum-corner coordinates, row then column."""

    eigen_probe: typing.Any = None
    eigen_weights: typing.Any = None
    algorithm_options: IterativeOptions = dataclasses.field(
        default_factory=RpieOptions)
    exitwave_options: typing.Optional[ExitWaveOptions] = None
    probe_options: typing.Optional[ProbeOptions] = None
    object_options: typing.Optional[ObjectOptions] = None
    position_options: typing.Optional[PositionOptions] = None

    def __post_init__(self):
        scan_shape = tuple(self.scan.shape)
        if (len(scan_shape) != 2 or scan_shape[1] != 2
                or any(n < 1 for n in scan_shape)):
            raise ValueError(f"scan shape {scan_shape} is incorrect. "
                             "It should be (N, 2) "
                             "where N >= 1 is the number of scan positions.")
        pshape = tuple(self.probe.shape)
        if (len(pshape) != 5 or pshape[:2] != (1, 1)
                or any(n < 1 for n in pshape) or pshape[-2] != pshape[-1]):
            raise ValueError(f"probe shape {pshape} is incorrect. "
                             "It should be (1, 1, S, W, H) "
                             "where S >=1 is the number of probes, and "
                             "W, H >= 1 are the square probe grid dimensions.")
        oshape = tuple(self.psi.shape)
        if len(oshape) != 3 or any(
                o <= p for o, p in zip(oshape[-2:], pshape[-2:])):
            raise ValueError(
                f"psi shape {oshape} is incorrect. "
                "It should be (D, W, H) where W, H > probe.shape[-2:].")
        check_allowed_positions(self.scan, self.psi, pshape)
        if self.exitwave_options is None:
            self.exitwave_options = ExitWaveOptions(
                measured_pixels=np.ones(pshape[-2:], dtype=np.bool_))

    def resample(self, factor: float, interp=None) -> "PtychoParameters":
        """Parameters rescaled by ``factor`` (options.py:170-195)."""
        interp = _resize_fft if interp is None else interp
        opt = lambda o, *a: o.resample(factor, *a) if o is not None else None
        return PtychoParameters(
            probe=interp(self.probe, factor),
            psi=_resize_spline(self.psi, factor),
            scan=self.scan * factor,
            eigen_probe=interp(self.eigen_probe, factor)
            if self.eigen_probe is not None else None,
            eigen_weights=self.eigen_weights,
            algorithm_options=self.algorithm_options,
            probe_options=opt(self.probe_options, interp),
            object_options=opt(self.object_options, interp),
            position_options=opt(self.position_options),
            exitwave_options=opt(self.exitwave_options),
        )

    def _map(self, arr, opts) -> "PtychoParameters":
        o = lambda x: getattr(x, opts)() if x is not None else None
        return PtychoParameters(
            probe=arr(self.probe, 'c64'),
            psi=arr(self.psi, 'c64'),
            scan=arr(self.scan, 'f32'),
            eigen_probe=arr(self.eigen_probe, 'c64'),
            eigen_weights=arr(self.eigen_weights, 'f32'),
            algorithm_options=self.algorithm_options,
            exitwave_options=o(self.exitwave_options),
            probe_options=o(self.probe_options),
            object_options=o(self.object_options),
            position_options=o(self.position_options),
        )

    def copy_to_device(self) -> "PtychoParameters":
        """Arrays as torch CUDA tensors on the current device."""
        return self._map(lambda x, d: to_device(x, dtype=d), 'copy_to_device')

    def copy_to_host(self) -> "PtychoParameters":
        """Arrays as NumPy arrays."""
        return self._map(lambda x, d: to_host(x), 'copy_to_host')

    @staticmethod
    def split(indices, *, x: "PtychoParameters") -> "PtychoParameters":
        """Parameters restricted to the positions ``indices``
        (options.py:266-290)."""
        c, f = precision.cfloating, precision.floating
        return PtychoParameters(
            probe=x.probe.astype(c),
            psi=x.psi.astype(c),
            scan=x.scan[indices].astype(f),
            eigen_probe=x.eigen_probe.astype(c)
            if x.eigen_probe is not None else None,
            eigen_weights=x.eigen_weights[indices].astype(f)
            if x.eigen_weights is not None else None,
            algorithm_options=copy.deepcopy(x.algorithm_options),
            exitwave_options=x.exitwave_options,
            probe_options=x.probe_options,
            object_options=x.object_options,
            position_options=x.position_options.split(indices)
            if x.position_options is not None else None,
        )

    @staticmethod
    def join(x, reorder, stripe_start) -> "PtychoParameters":
        """Recombine per-worker parameters (options.py:292-330)."""
        return PtychoParameters(
            probe=x[0].probe,
            psi=ObjectOptions.join_psi([e.psi for e in x],
                                       probe_width=x[0].probe.shape[-2],
                                       stripe_start=stripe_start),
            scan=np.concatenate([e.scan for e in x], axis=0)[reorder],
            eigen_probe=x[0].eigen_probe,
            eigen_weights=np.concatenate([e.eigen_weights for e in x],
                                         axis=0)[reorder]
            if x[0].eigen_weights is not None else None,
            algorithm_options=x[0].algorithm_options,
            exitwave_options=x[0].exitwave_options,
            probe_options=x[0].probe_options,
            object_options=ObjectOptions.join(
                [e.object_options for e in x], stripe_start=stripe_start,
                probe_width=x[0].probe.shape[-2])
            if x[0].object_options is not None else None,
            position_options=PositionOptions.join(
                [e.position_options for e in x], reorder),
        )


def _resize_spline(x: np.ndarray, f: float) -> np.ndarray:
    return scipy.ndimage.zoom(x, zoom=[1] * (x.ndim - 2) + [f, f],
                              grid_mode=True, prefilter=False)


def _resize_cv(x: np.ndarray, f: float, interpolation: int) -> np.ndarray:
    """OpenCV interpolation of the last two axes, real and imaginary parts
    separately (options.py:342-353 -> view.resize_complex_image)."""
    import cv2
    lead = x.shape[:-2]
    size = (int(x.shape[-1] * f), int(x.shape[-2] * f))  # cv2 wants (width, height)
    planes = []
    for img in x.reshape(-1, *x.shape[-2:]):
        re = cv2.resize(np.ascontiguousarray(img.real), size, interpolation=interpolation)
        im = cv2.resize(np.ascontiguousarray(img.imag), size, interpolation=interpolation)
        planes.append(re + 1j * im)
    out = np.asarray(planes)
    return out.reshape(*lead, *out.shape[-2:])


def _resize_linear(x: np.ndarray, f: float) -> np.ndarray:
    return _resize_cv(x, f, 1)  # cv2.INTER_LINEAR


def _resize_cubic(x: np.ndarray, f: float) -> np.ndarray:
    return _resize_cv(x, f, 2)  # cv2.INTER_CUBIC


def _resize_lanczos(x: np.ndarray, f: float) -> np.ndarray:
    return _resize_cv(x, f, 4)  # cv2.INTER_LANCZOS4


def pad_fourier_space(x: np.ndarray, w: int) -> np.ndarray:
    """Zero-pad a DC-at-corner spectrum to w x w (options.py:382-391)."""
    assert x.shape[-2] == x.shape[-1], "Only works on square arrays right now."
    half1 = x.shape[-1] // 2
    half0 = x.shape[-1] - half1
    cols = np.r_[0:half0, (w - half1):w]
    out = np.zeros_like(x, shape=(*x.shape[:-2], w, w))
    out[..., 0:half0, cols] = x[..., 0:half0, :]
    out[..., -half1:w, cols] = x[..., -half1:, :]
    return out


def _resize_fft(x: np.ndarray, f: float) -> np.ndarray:
    """Fourier resampling of the last two axes (options.py:393-409)."""
    if f == 1:
        return x
    crop_or_pad = crop_fourier_space if f < 1 else pad_fourier_space
    return np.fft.ifft2(
        crop_or_pad(np.fft.fft2(x, norm='ortho', axes=(-2, -1)),
                    w=int(x.shape[-1] * f)),
        norm='ortho', axes=(-2, -1))
