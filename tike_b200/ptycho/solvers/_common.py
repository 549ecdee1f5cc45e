"""Shared host-side helpers of the solvers: batch staging, masks, multi-GPU
reductions.  (No reference counterpart file; replaces the glue spread over
communicators/stream.py:285-404 and the top of each solver.)"""
from __future__ import annotations

import os as _os

import numpy as np
import torch

from ..._array import to_host

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


def _staged_dtype(data):
    """float32, or uint16 when the patterns are 16-bit counts (ptycho.py:387-389)."""
    dt = data.dtype
    if dt in (torch.uint16, np.dtype(np.uint16)):
        return torch.uint16
    return torch.float32


def _host_slice(data, lo, hi, dtype):
    """data[lo:hi] as a host tensor of ``dtype`` (no copy for pinned tensors)."""
    if isinstance(data, torch.Tensor):
        part = data[lo:hi]
        return part if part.dtype == dtype else part.to(dtype)
    host = np.asarray(data[lo:hi])
    want = np.uint16 if dtype == torch.uint16 else np.float32
    return torch.from_numpy(np.ascontiguousarray(host, dtype=want))


class BatchStager:
    """Delivers the diffraction patterns of each batch as device tensors.

    Device-resident data is sliced.  Host (pinned) data is uploaded on a side
    stream in sub-batch chunks, ``depth`` chunks ahead of the compute stream
    and across batch boundaries, so the H2D copy of chunk j+1 overlaps the
    kernels of chunk j (the reference triple-buffers 64-pattern chunks the
    same way, stream.py:359-404)."""

    def __init__(self, data, batches, sequence, device, chunk_positions=None,
                 depth=2, cuts=None):
        """``cuts`` (optional, one absolute index per batch of ``batches``):
        a piece never straddles the cut of its batch, so a solver can start the
        inter-GPU exchange once the positions before the cut are done."""
        self.data, self.batches, self.sequence = data, batches, list(sequence)
        self.device = device
        self.resident = (isinstance(data, torch.Tensor) and data.is_cuda) or (
            not isinstance(data, (torch.Tensor, np.ndarray))
            and hasattr(data, '__cuda_array_interface__'))
        self.depth = max(1, int(depth))
        if chunk_positions is None:
            # pieces of about 512 MiB of float32 patterns (8192 at 128 x 128): few
            # enough launches per batch that their ramp-up does not show, small
            # enough that the first piece of a batch is on the device in time
            per_pattern = int(np.prod(data.shape[1:])) * 4
            default = max(256, min(8192, (512 << 20) // max(per_pattern, 1)))
            chunk_positions = int(_os.environ.get('TB_STAGE_CHUNK', default))
        # flat list of (k, lo, hi) over the whole epoch
        self._plan, self._first = [], {}
        for k in range(len(self.sequence)):
            lo, hi = self._range(k)
            self._first[k] = len(self._plan)
            step = (hi - lo) if (self.resident or chunk_positions <= 0) else chunk_positions
            cut = int(cuts[self.sequence[k]]) if cuts is not None else lo
            c = lo
            while c < hi or (c == lo and hi == lo):
                end = min(hi, c + max(step, 1))
                if c < cut < end:
                    end = cut
                self._plan.append((k, c, end))
                c = end
                if hi == lo:
                    break
        self._pending = {}
        self._consumed = {}
        self._stream = None if self.resident else torch.cuda.Stream(device=device)
        if not self.resident:
            # ring of depth + 1 fixed device buffers: no allocator traffic while
            # the epoch runs, reuse ordered by events
            rows = max((hi - lo for _, lo, hi in self._plan), default=0)
            dt = _staged_dtype(data)
            self._ring = [torch.empty((rows, *tuple(data.shape[1:])), dtype=dt, device=device)
                          for _ in range(self.depth + 1)]
            for j in range(min(self.depth, len(self._plan))):
                self._issue(j)

    def _range(self, k):
        b = self.batches[self.sequence[k]]
        return int(b[0]), int(b[-1]) + 1

    def _issue(self, j):
        if j in self._pending or j >= len(self._plan):
            return
        _, lo, hi = self._plan[j]
        slot = j % len(self._ring)
        with torch.cuda.stream(self._stream):
            prev = self._consumed.pop(j - len(self._ring), None)
            if prev is not None:
                self._stream.wait_event(prev)  # the kernels that read this slot
            chunk = self._ring[slot][:hi - lo]
            chunk.copy_(_host_slice(self.data, lo, hi, chunk.dtype), non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._stream)
        self._pending[j] = (chunk, done)

    def chunks(self, k):
        """Yield ``(lo, hi, patterns)`` covering the k-th batch of the sequence.
        A yielded piece is valid until the next one is requested."""
        j = self._first[k]
        while j < len(self._plan) and self._plan[j][0] == k:
            _, lo, hi = self._plan[j]
            if self.resident:
                yield lo, hi, stage_data(self.data, lo, hi, self.device)
            else:
                self._issue(j)
                chunk, done = self._pending.pop(j)
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(done)
                for ahead in range(1, self.depth):
                    self._issue(j + ahead)
                yield lo, hi, chunk
                used = torch.cuda.Event()
                used.record(torch.cuda.current_stream(self.device))
                self._consumed[j] = used
                self._issue(j + self.depth)
            j += 1

    def get(self, k):
        """Patterns of the whole k-th batch of the sequence (one device tensor)."""
        if self.resident:
            parts = [c for _, _, c in self.chunks(k)]
            return parts[0] if len(parts) == 1 else torch.cat(parts)
        return torch.cat([c.clone() for _, _, c in self.chunks(k)])


def own_costs(costs, worker_index: int, memory_length: int = 3):
    """This worker's last ``memory_length`` epoch costs for the checked
    momentum (lstsq.py:255-262).  Finished epochs hold one cost per rank; the
    row of the running epoch only holds the local cost (the reference indexes
    it with worker_index as well and raises IndexError for workers > 0)."""
    return [float(x[worker_index] if worker_index < len(x) else x[-1])
            for x in costs[-memory_length:]]


def detector_width(data) -> int:
    return int(data.shape[-1])


def precond_max_of(preconditioner):
    """Cross-rank per-slice max(Re preconditioner) attached by
    update_preconditioners when the object rows are split over ranks, else None
    (the update kernels then take the maximum of the local array)."""
    return getattr(preconditioner, '_tb_max', None)


class ObjectReducer:
    """Sum of object-sized (D, H, W) arrays over the ranks.

    Without a row plan (``comm.plan is None``) it is the NCCL all-reduce of
    the whole array.  With one (communicators.RowPlan) only the rows that two
    ranks touch are exchanged (Comm.halo_sum_), and the exchange can run on a
    side stream while this rank's interior positions -- the ones whose
    footprint touches no shared row -- are still being processed:

        red.begin(t)     # after the last boundary position of the batch
        ...              # more kernels accumulating into unshared rows of t
        red.finish(t)    # before t is consumed

    (north_star: "NCCL ... overlapped with the next batch"; SURVEY 8e: the
    exchange may hide behind the scatter kernel's progress, not behind the
    next batch, because the next batch needs the updated object.)"""

    _SIDE = {}

    def __init__(self, comm):
        self.comm = comm if (comm is not None and comm.size > 1) else None
        self.plan = getattr(comm, 'plan', None) if self.comm is not None else None
        self._pending = {}

    @property
    def active(self):
        return self.comm is not None and not _SKIP_ALLREDUCE

    def _side(self, device):
        key = str(device)
        if key not in ObjectReducer._SIDE:
            ObjectReducer._SIDE[key] = torch.cuda.Stream(device=device)
        return ObjectReducer._SIDE[key]

    def begin(self, t):
        if not self.active or t is None or self.plan is None or not t.is_cuda:
            return
        if id(t) in self._pending:
            return
        cur = torch.cuda.current_stream(t.device)
        side = self._side(t.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            self.comm.halo_sum_(t, self.plan)
            done = torch.cuda.Event()
            done.record(side)
        self._pending[id(t)] = done

    def finish(self, t):
        if not self.active or t is None:
            return
        if self.plan is None:
            self.comm.allreduce_sum_(t)
            return
        done = self._pending.pop(id(t), None)
        if done is None:
            self.comm.halo_sum_(t, self.plan)
        else:
            torch.cuda.current_stream(t.device).wait_event(done)


def allreduce_(comm, *tensors):
    """Sum tensors over all ranks in place (no-op without a communicator)."""
    if comm is None or comm.size == 1:
        return
    if _SKIP_ALLREDUCE:  # development switch: isolate the cost of the collectives
        return
    for t in tensors:
        if t is not None:
            comm.allreduce_sum_(t)
