"""Optimisation helpers used by the solvers (reference: src/tike/opt.py)."""
from __future__ import annotations

import numpy as np

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _zeros_like_real(g):
    if torch is not None and isinstance(g, torch.Tensor):
        return torch.zeros_like(g.real)
    return np.zeros_like(g.real)


def _zeros_like(g):
    if torch is not None and isinstance(g, torch.Tensor):
        return torch.zeros_like(g)
    return np.zeros_like(g)


def _sqrt(x):
    if torch is not None and isinstance(x, torch.Tensor):
        return torch.sqrt(x)
    return np.sqrt(x)


def momentum(g, v, m, vdecay=None, mdecay=0.9):
    """First-moment smoothing (opt.py:67-82): returns (m, None, m)."""
    m = 0 if m is None else m
    m = mdecay * m + (1 - mdecay) * g
    return m, None, m


def adam(g, v=None, m=None, vdecay=0.999, mdecay=0.9, eps=1e-8):
    """ADAM direction as written in the reference (opt.py:165-213): moments
    are divided by (1 - decay) without the usual power of the step count."""
    v = _zeros_like_real(g) if v is None else v
    m = _zeros_like(g) if m is None else m
    m = mdecay * m + (1 - mdecay) * g
    v = vdecay * v + (1 - vdecay) * (g * g.conj()).real
    m_hat = m / (1 - mdecay)
    v_hat = _sqrt(v / (1 - vdecay))
    return m_hat / (v_hat + eps), v, m


def fit_line_least_squares(y, x):
    """(slope, intercept) of the least-squares line y = slope * x + intercept
    (same estimator as the reference's opt.py:383-400, centred form)."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if x.size != y.size or x.size == 0:
        raise ValueError('x and y must be non-empty and of equal length')
    xm, ym = x.mean(), y.mean()
    dx = x - xm
    slope = float(np.dot(dx, y - ym) / np.dot(dx, dx))
    return slope, float(ym - slope * xm)
