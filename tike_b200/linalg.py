"""Small linear-algebra helpers that work on NumPy arrays and torch tensors
(reference: src/tike/linalg.py:12-137)."""
from __future__ import annotations

import numpy as np

try:  # torch is plumbing for device arrays; NumPy inputs stay NumPy
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _abs2(x):
    if _is_torch(x):
        return (x * x.conj()).real if x.is_complex() else x * x
    return (x * np.conj(x)).real


def _dims(axis):
    if axis is None or isinstance(axis, int):
        return axis
    return tuple(axis)


def mnorm(x, axis=None, keepdims=False):
    """sqrt(mean(|x|^2)) (linalg.py:12-14)."""
    a = _abs2(x)
    if _is_torch(x):
        m = a.mean() if axis is None else a.mean(dim=_dims(axis), keepdim=keepdims)
        return torch.sqrt(m)
    return np.sqrt(np.mean(a, axis=_dims(axis), keepdims=keepdims))


def norm(x, axis=None, keepdims=False):
    """sqrt(sum(|x|^2)) (linalg.py:17-19)."""
    a = _abs2(x)
    if _is_torch(x):
        s = a.sum() if axis is None else a.sum(dim=_dims(axis), keepdim=keepdims)
        return torch.sqrt(s)
    return np.sqrt(np.sum(a, axis=_dims(axis), keepdims=keepdims))


def inner(x, y, axis=None, keepdims=False):
    """sum(x * conj(y)) (linalg.py:28-30)."""
    p = x * y.conj()
    if _is_torch(p):
        return p.sum() if axis is None else p.sum(dim=_dims(axis), keepdim=keepdims)
    return p.sum(axis=_dims(axis), keepdims=keepdims)


def projection(a, b, axis=None):
    """Complex projection of a onto b (linalg.py:22-25)."""
    bh = b / inner(b, b, axis=axis, keepdims=True)
    return inner(a, b, axis=axis, keepdims=True) * bh


def orthogonalize_gs(x, axis=-1, N=None):
    """Gram-Schmidt over axis N with inner products over ``axis``
    (linalg.py:67-108)."""
    ndim = x.ndim
    try:
        axis = tuple(a % ndim for a in axis)
    except TypeError:
        axis = (axis % ndim,)
    if N is None:
        N = ndim - 1
        while N in axis:
            N -= 1
    N = N % ndim
    if N in axis:
        raise ValueError("Cannot orthogonalize a single vector.")
    if _is_torch(x):
        x = torch.movedim(x, N, 0)
        u = x.clone()
        for i in range(1, len(x)):
            u[i:] -= projection(x[i:], u[i - 1:i], axis=axis)
        return torch.movedim(u, 0, N)
    x = np.moveaxis(x, N, 0)
    u = x.copy()
    for i in range(1, len(x)):
        u[i:] -= projection(x[i:], u[i - 1:i], axis=axis)
    return np.moveaxis(u, 0, N)


def hermitian(x):
    """Conjugate transpose of the last two axes (linalg.py:103-105)."""
    if _is_torch(x):
        return x.conj().transpose(-1, -2)
    return np.conj(x).swapaxes(-1, -2)


def lstsq(a, b, weights=None):
    """Weighted least squares through the normal equations,
    inv(a^H W a) a^H W b, batched over leading axes (linalg.py:33-58)."""
    if weights is not None:
        root = weights**0.5
        a = a * root[..., None]
        b = b * root[..., None]
    ah = hermitian(a)
    if _is_torch(a):
        return torch.linalg.inv(ah @ a) @ ah @ b
    return np.linalg.inv(ah @ a) @ ah @ b


def cov(x):
    """Scatter matrix of observations along axis -2 (linalg.py:108-111)."""
    centred = x - x.mean(-2, keepdims=True) if not _is_torch(x) else x - x.mean(-2, keepdim=True)
    return hermitian(centred) @ centred


def pca_eig(data, k: int):
    """The k leading principal components of (..., N, D) data through an
    eigen-decomposition of the scatter matrix; returns (S, U), largest first
    (linalg.py:114-137)."""
    c = cov(data)
    if _is_torch(c):
        S, U = torch.linalg.eigh(c)
        return torch.flip(S[..., -k:], dims=(-1,)), torch.flip(U[..., -k:], dims=(-1,))
    S, U = np.linalg.eigh(c)
    return S[..., ::-1][..., :k], U[..., ::-1][..., :k]
