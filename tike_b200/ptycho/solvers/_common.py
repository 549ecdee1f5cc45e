"""Shared host-side helpers of the solvers: batch staging, masks, multi-GPU
reductions.  (No reference counterpart file; replaces the glue spread over
communicators/stream.py:285-404 and the top of each solver.)"""
from __future__ import annotations

import numpy as np
import torch

from ... import kernels
from ..._array import to_device, to_host


import os as _os

_SKIP_ALLREDUCE = bool(_os.environ.get('TB_DEBUG_SKIP_ALLREDUCE'))


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


class BatchStager:
    """Delivers the diffraction patterns of each batch as a device tensor.

    Device-resident data is sliced.  Host (pinned) data is uploaded on a side
    stream one batch ahead of the compute stream, so the H2D copy of batch
    k+1 overlaps the kernels of batch k (the reference triple-buffers
    64-pattern chunks the same way, stream.py:359-404)."""

    def __init__(self, data, batches, sequence, device):
        self.data, self.batches, self.sequence = data, batches, list(sequence)
        self.device = device
        self.resident = (isinstance(data, torch.Tensor) and data.is_cuda) or (
            not isinstance(data, (torch.Tensor, np.ndarray))
            and hasattr(data, '__cuda_array_interface__'))
        self._pending = {}
        self._stream = None if self.resident else torch.cuda.Stream(device=device)
        if not self.resident and self.sequence:
            self._prefetch(0)

    def _range(self, k):
        b = self.batches[self.sequence[k]]
        return int(b[0]), int(b[-1]) + 1

    def _prefetch(self, k):
        lo, hi = self._range(k)
        with torch.cuda.stream(self._stream):
            chunk = stage_data(self.data, lo, hi, self.device)
            done = torch.cuda.Event()
            done.record(self._stream)
        self._pending[k] = (chunk, done)

    def get(self, k):
        """Patterns of the k-th batch of the sequence (device tensor)."""
        lo, hi = self._range(k)
        if self.resident:
            return stage_data(self.data, lo, hi, self.device)
        if k not in self._pending:
            self._prefetch(k)
        chunk, done = self._pending.pop(k)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(done)
        chunk.record_stream(cur)
        if k + 1 < len(self.sequence):
            self._prefetch(k + 1)
        return chunk


def detector_width(data) -> int:
    return int(data.shape[-1])


def allreduce_(comm, *tensors):
    """Sum tensors over all ranks in place (no-op without a communicator)."""
    if comm is None or comm.size == 1:
        return
    if _SKIP_ALLREDUCE:  # development switch: isolate the cost of the collectives
        return
    for t in tensors:
        if t is not None:
            comm.allreduce_sum_(t)
