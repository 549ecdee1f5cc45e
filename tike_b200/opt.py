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


# ---------------------------------------------------------------------------
# Utilities of the reference's opt module that the current solvers do not call
# but user scripts do (opt.py:21-64, 85-162).  The conjugate-gradient machinery
# (line_search, direction_dy, conjugate_gradient, ...) belongs to solvers that
# no longer exist in the reference and is not carried over.
# ---------------------------------------------------------------------------

def is_converged(algorithm_options) -> bool:
    """True when the least-squares slope of the last ``convergence_window``
    epoch costs is not negative, tested every half window (opt.py:21-43)."""
    window = algorithm_options.convergence_window
    costs = algorithm_options.costs
    if window < 2 or len(costs) < window or len(costs) % window // 2 != 0:
        return False
    recent = np.asarray(costs[-window:], dtype=float).reshape(window, -1).mean(axis=1)
    slope, _ = fit_line_least_squares(y=recent, x=np.arange(window))
    return slope >= 0


def batch_indicies(n, m=1, use_random=True):
    """The indices [0, n) as m nearly equal groups (opt.py:46-54)."""
    if not 0 < m <= n:
        raise AssertionError((m, n))
    from . import random as tb_random
    order = tb_random.randomizer_np.permutation(n) if use_random else np.arange(n)
    return np.array_split(order, m)


def get_batch(x, b, n):
    """x[b[n]] (opt.py:57-59)."""
    return x[b[n]]


def put_batch(y, x, b, n):
    """x[b[n]] = y (opt.py:62-64)."""
    x[b[n]] = y


def adagrad(g, v=None, m=None, eps=1e-6):
    """AdaGrad direction (opt.py:85-122): the first call only seeds the
    squared-gradient sum and returns the gradient itself."""
    if v is None:
        return g, (g * g.conj()).real, m
    v = v + (g * g.conj()).real
    return g / _sqrt(v + eps), v, m


def adadelta(g, d0=None, v=None, m=None, decay=0.9, eps=1e-6):
    """AdaDelta direction (opt.py:125-162)."""
    v = 0 if v is None else v
    m = 0 if m is None else m
    d0 = 0 if d0 is None else d0
    d0_abs2 = d0 * d0 if isinstance(d0, (int, float)) else (d0 * d0.conj()).real
    v = v * decay + (1 - decay) * (g * g.conj()).real
    m = m * decay + (1 - decay) * d0_abs2
    return _sqrt((m + eps) / (v + eps)) * g, v, m
