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


# ---------------------------------------------------------------------------
# Poisson step lengths as array functions (exitwave.py:122-234).  The solvers
# evaluate the same two fixed-point iterations inside the fused kernels
# (csrc/rpie_fast.cu, rpie.cu, large.cu); these versions serve user code and
# work on NumPy arrays or device tensors alike.
# ---------------------------------------------------------------------------

def _masked_sum(x, measured_pixels):
    """Sum over the detector pixels selected by the boolean (W, H) mask."""
    return x[..., measured_pixels].sum(-1)


def poisson_steplength_all_modes(xi, abs2_Psi, I_e, I_m, measured_pixels,
                                 step_length, weight_avg):
    """One step length per exit-wave mode.

    xi (F, 1, 1, W, H) = 1 - I_m / I_e; abs2_Psi (F, 1, S, W, H); I_e, I_m
    (F, W, H); step_length (F, 1, S, 1, 1).  Two damped fixed-point updates
    alpha <- (1 - w) alpha + w * sum(xi |Psi|^2 (1 + I_m (xi alpha - 1) / D))
    / sum(xi^2 |Psi|^2), D = |Psi|^2 (xi alpha - 1)^2 + I_e - |Psi|^2."""
    I_e = I_e[:, None, None, ...]
    I_m = I_m[:, None, None, ...]
    weighted = xi * abs2_Psi
    normaliser = _masked_sum(xi * weighted, measured_pixels)
    for _ in range(2):
        t = xi * step_length - 1
        spread = abs2_Psi * t * t + I_e - abs2_Psi
        numerator = _masked_sum(weighted * (1 + (I_m * t) / spread), measured_pixels)
        step_length = (step_length * (1 - weight_avg) +
                       (numerator / normaliser)[..., None, None] * weight_avg)
    return step_length


def poisson_steplength_dominant_mode(xi, I_e, I_m, measured_pixels, step_length,
                                     weight_avg):
    """One step length for all modes, from the total intensity only
    (exitwave.py:183-234); shapes as in poisson_steplength_all_modes."""
    I_e = I_e[:, None, None, ...]
    I_m = I_m[:, None, None, ...]
    normaliser = _masked_sum(xi * xi * I_e, measured_pixels)
    for _ in range(2):
        numerator = _masked_sum(xi * (I_e - I_m / (1 - step_length * xi)), measured_pixels)
        step_length = ((1 - weight_avg) * step_length +
                       weight_avg * (numerator / normaliser)[..., None, None])
    return step_length
