"""Seeded synthetic ptychography inputs (NumPy, host side).

Used by the tests, the golden-vector generator and ``bench.py``.  The recipes
follow SURVEY.md §8(d): smooth random complex object, Gaussian-windowed disc
probe with quadratic phase and Hermite-like higher modes, raster scan with
uniform jitter kept inside ``check_allowed_positions`` bounds
(reference: src/tike/ptycho/position.py:600-628).
"""
from __future__ import annotations

import numpy as np

__all__ = ['make_object', 'make_probe', 'make_scan', 'make_problem']


def _smooth_field(shape, rng, cutoff=0.05):
    """Low-pass filtered white noise rescaled to [0, 1]."""
    noise = rng.standard_normal(shape).astype(np.float32)
    fy = np.fft.fftfreq(shape[0])[:, None]
    fx = np.fft.fftfreq(shape[1])[None, :]
    lp = np.exp(-(fy * fy + fx * fx) / (2 * cutoff * cutoff))
    field = np.fft.ifft2(np.fft.fft2(noise) * lp).real
    field -= field.min()
    field /= max(field.max(), 1e-12)
    return field.astype(np.float32)


def make_object(height: int, width: int, seed: int = 0) -> np.ndarray:
    """(1, H, W) complex64: amplitude 0.8+0.2u, phase pi*(v-0.5)."""
    rng = np.random.default_rng(seed)
    u = _smooth_field((height, width), rng)
    v = _smooth_field((height, width), rng)
    psi = (0.8 + 0.2 * u) * np.exp(1j * np.pi * (v - 0.5))
    return psi.astype(np.complex64)[None]


def make_probe(width: int, nmodes: int = 1, seed: int = 2,
               photons: float = 1.0) -> np.ndarray:
    """(1, 1, M, N, N) complex64 probe with M mutually different modes."""
    rng = np.random.default_rng(seed)
    c = (np.arange(width, dtype=np.float32) + 0.5) / width - 0.5
    y, x = np.meshgrid(c, c, indexing='ij')
    r2 = x * x + y * y
    base = np.exp(-r2 / (2 * 0.18**2)) * np.exp(1j * 40.0 * r2)
    base = base * (r2 < 0.45**2)
    modes = []
    for m in range(nmodes):
        i, j = divmod(m, 3)
        poly = (4 * x)**j * (4 * y)**i
        phase = np.exp(2j * np.pi * rng.random())
        mode = base * poly * phase
        mode = mode / np.sqrt(np.sum(np.abs(mode)**2)) / (m + 1)
        modes.append(mode)
    probe = np.stack(modes, axis=0)[None, None] * np.sqrt(photons)
    return probe.astype(np.complex64)


def make_scan(npos: int, height: int, width: int, probe_width: int,
              seed: int = 1, jitter: float = 2.0,
              margin: float = 0.0) -> np.ndarray:
    """(P, 2) float32 raster scan + uniform jitter, kept >= 1 and
    <= dim - probe_width - 1 - eps so floor(scan) is an allowed corner."""
    rng = np.random.default_rng(seed)
    lo = 1.0 + jitter + margin
    hi_y = height - probe_width - 2.0 - jitter - margin
    hi_x = width - probe_width - 2.0 - jitter - margin
    if hi_y <= lo or hi_x <= lo:
        raise ValueError('object too small for this probe/jitter')
    aspect = (hi_x - lo) / (hi_y - lo)
    ny = max(1, int(np.ceil(np.sqrt(npos / aspect))))
    nx = max(1, int(np.ceil(npos / ny)))
    gy = np.linspace(lo, hi_y, ny, dtype=np.float64)
    gx = np.linspace(lo, hi_x, nx, dtype=np.float64)
    yy, xx = np.meshgrid(gy, gx, indexing='ij')
    grid = np.stack([yy.ravel(), xx.ravel()], axis=1)[:npos]
    grid = grid + rng.uniform(-jitter, jitter, size=grid.shape)
    grid[:, 0] = np.clip(grid[:, 0], 1.0 + margin,
                         height - probe_width - 1.001 - margin)
    grid[:, 1] = np.clip(grid[:, 1], 1.0 + margin,
                         width - probe_width - 1.001 - margin)
    return grid.astype(np.float32)


def make_problem(npos: int, probe_width: int, nmodes: int, height: int,
                 width: int, seed: int = 0, margin: float = 0.0):
    """Return (psi_true, probe, scan) for a synthetic experiment."""
    psi = make_object(height, width, seed)
    probe = make_probe(probe_width, nmodes, seed + 2,
                       photons=float(probe_width * probe_width) * 50.0)
    scan = make_scan(npos, height, width, probe_width, seed + 1,
                     margin=margin)
    return psi, probe, scan
