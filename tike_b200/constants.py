"""Physical constants and small wave helpers (reference: src/tike/constants.py).

Energies in keV, lengths in cm, as in the reference."""
import numpy as np

__all__ = [
    'PLANCK_CONSTANT', 'SPEED_OF_LIGHT', 'wavelength', 'wavenumber',
    'complex_amplitude', 'complex_intensity', 'complex_phase', 'sum_square_norm',
]

PLANCK_CONSTANT = 6.58211928e-19  # reduced Planck constant [keV s]
SPEED_OF_LIGHT = 299792458e+2  # [cm / s]


def wavelength(energy):
    """Wavelength [cm] of photons of ``energy`` [keV]: 2 pi hbar c / E."""
    return 2 * np.pi * PLANCK_CONSTANT * SPEED_OF_LIGHT / energy


def wavenumber(energy):
    """Wavenumber [1 / cm] of photons of ``energy`` [keV]: E / (hbar c)."""
    return energy / (PLANCK_CONSTANT * SPEED_OF_LIGHT)


def complex_amplitude(probe_grid):
    return np.abs(probe_grid)


def complex_intensity(probe_grid):
    return np.abs(probe_grid)**2


def complex_phase(probe_grid):
    return np.angle(probe_grid)


def sum_square_norm(x, N=1):
    """``x`` rescaled so that the sum of its squared magnitudes equals N."""
    return x * np.sqrt(N / np.sum(np.abs(x)**2))
