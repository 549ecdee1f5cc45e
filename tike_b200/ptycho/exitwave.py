"""Exit-wave (detector plane) options (reference: src/tike/ptycho/exitwave.py:22-120)."""
from __future__ import annotations

import dataclasses

import numpy as np

from .._array import to_device, to_host


def crop_fourier_space(x, w: int):
    """Keep the w x w lowest frequencies of a DC-at-corner array
    (exitwave.py:237-248, solvers/options.py:368-380)."""
    assert x.shape[-2] == x.shape[-1], "Only works on square arrays right now."
    half1 = w // 2
    half0 = w - half1
    keep = np.r_[0:half0, (x.shape[-1] - half1):x.shape[-1]]
    return x[..., keep][..., keep, :]


@dataclasses.dataclass
class ExitWaveOptions:
    """Settings of the far-field update (same fields as the reference)."""

    measured_pixels: np.ndarray
    """Boolean (detector, detector) mask: True where the detector measured."""

    noise_model: str = "gaussian"
    """'gaussian' or 'poisson'."""

    step_length_weight: float = 0.5
    step_length_usemodes: str = "all_modes"
    step_length_start: float = 0.5
    unmeasured_pixels_scaling: float = 1.00
    propagation_normalization: str = "ortho"

    def copy_to_device(self) -> "ExitWaveOptions":
        return dataclasses.replace(
            self, measured_pixels=to_device(self.measured_pixels, dtype='bool'))

    def copy_to_host(self) -> "ExitWaveOptions":
        return dataclasses.replace(
            self, measured_pixels=to_host(self.measured_pixels))

    def resample(self, factor: float) -> "ExitWaveOptions":
        mask = to_host(self.measured_pixels)
        return dataclasses.replace(
            self,
            measured_pixels=crop_fourier_space(
                mask, int(mask.shape[-1] * factor)),
        )
