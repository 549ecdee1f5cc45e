"""Stub of cupyx for the NumPy-backed shim (test infrastructure only)."""
import numpy as _np

from . import scipy  # noqa: F401


def empty_pinned(shape, dtype=float, order='C'):
    return _np.empty(shape, dtype=dtype, order=order)


def zeros_pinned(shape, dtype=float, order='C'):
    return _np.zeros(shape, dtype=dtype, order=order)
