"""Shared host-side helpers of the solvers: batch staging, masks, multi-GPU
reductions.  (No reference counterpart file; replaces the glue spread over
communicators/stream.py:285-404 and the top of each solver.)"""
from __future__ import annotations

import numpy as np
import torch

from ... import kernels
from ..._array import to_device, to_host


class MaskInfo:
    """Device uint8 copy of ExitWaveOptions.measured_pixels plus its count."""

    def __init__(self, measured_pixels, device):
        host = np.asarray(to_host(measured_pixels)).astype(bool)
        self.count = int(host.sum())
        if self.count <= 0:
            raise ValueError('measured_pixels must contain at least one True')
        self.all = bool(host.all())
        self.shape = host.shape
        self.dev = None if self.all else torch.as_tensor(
            host.astype(np.uint8)).to(device).contiguous()


def stage_data(data, lo: int, hi: int, device):
    """Return data[lo:hi] as a device tensor (float32 or uint16).

    Resident device arrays are sliced (no copy).  Host arrays are uploaded on
    the current stream — the reference re-streams every chunk every epoch
    (stream.py:380-404); keep the data resident to avoid that."""
    if isinstance(data, torch.Tensor):
        chunk = data[lo:hi]
        if chunk.is_cuda:
            return chunk.contiguous()
        if chunk.dtype not in (torch.float32, torch.uint16):
            chunk = chunk.to(torch.float32)
        return chunk.to(device, non_blocking=True)
    if hasattr(data, '__cuda_array_interface__'):
        return torch.as_tensor(data, device='cuda')[lo:hi].contiguous()
    host = np.asarray(data[lo:hi])
    if host.dtype == np.uint16:
        t = torch.from_numpy(np.ascontiguousarray(host))
    else:
        t = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32))
    return t.to(device, non_blocking=True)


def detector_width(data) -> int:
    return int(data.shape[-1])


def allreduce_(comm, *tensors):
    """Sum tensors over all ranks in place (no-op without a communicator)."""
    if comm is None or comm.size == 1:
        return
    for t in tensors:
        if t is not None:
            comm.allreduce_sum_(t)
