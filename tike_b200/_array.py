"""Host <-> device array plumbing (torch owns device memory).

The reference keeps parameters as NumPy arrays on the host and CuPy arrays on
the device (options.py:197-264).  Here device arrays are torch CUDA tensors;
any ``__cuda_array_interface__`` exporter handed in by a caller (e.g. CuPy) is
viewed zero-copy.
"""
from __future__ import annotations

import numpy as np
import torch

_TORCH_DTYPES = {
    'c64': torch.complex64,
    'f32': torch.float32,
    'bool': torch.bool,
    'u8': torch.uint8,
}


def is_device(x) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda) or (
        not isinstance(x, (np.ndarray, torch.Tensor))
        and hasattr(x, '__cuda_array_interface__'))


def to_device(x, dtype=None, device=None):
    """Return a contiguous torch CUDA tensor (None passes through)."""
    if x is None:
        return None
    if isinstance(dtype, str):
        dtype = _TORCH_DTYPES[dtype]
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, '__cuda_array_interface__'):
        t = torch.as_tensor(x, device='cuda')
    else:
        a = np.asarray(x)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.complex128:
            a = a.astype(np.complex64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    if not t.is_cuda or (t.device != torch.device(device)):
        t = t.to(device)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def to_host(x):
    """Return a NumPy array (None passes through)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if hasattr(x, '__cuda_array_interface__'):
        return torch.as_tensor(x, device='cuda').cpu().numpy()
    return np.asarray(x)


def pinned(x: np.ndarray) -> torch.Tensor:
    """Copy a host array into page-locked memory (reference:
    cluster._split_pinned, cluster.py:33-41)."""
    t = torch.from_numpy(np.ascontiguousarray(x))
    if torch.cuda.is_available():
        try:
            return t.pin_memory()
        except RuntimeError:  # pragma: no cover
            return t
    return t
