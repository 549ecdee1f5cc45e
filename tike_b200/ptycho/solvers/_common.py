"""Shared host-side helpers of the solvers: batch staging, masks, multi-GPU
reductions.  (No reference counterpart file; replaces the glue spread over
communicators/stream.py:285-404 and the top of each solver.)"""
from __future__ import annotations

import os as _os
import weakref as _weakref

import numpy as np
import torch

from ... import random as tb_random
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


class _HostRing:
    """Pinned-host -> device upload ring of one data array, kept ACROSS epochs.

    ``depth + 1`` fixed device buffers and one side stream.  ``issue(lo, hi)``
    starts the copy of ``data[lo:hi]`` into the next slot (after the kernels
    that read that slot, ordered by events); ``take`` hands the piece to the
    compute stream.  Keeping the ring alive between solver calls lets the
    last pieces of an epoch overlap the first copies of the next one
    (BatchStager predicts them from the next batch order)."""

    def __init__(self, data, rows, device, depth):
        self.data, self.device, self.rows = data, device, rows
        dt = _staged_dtype(data)
        self.buffers = [torch.empty((rows, *tuple(data.shape[1:])), dtype=dt, device=device)
                        for _ in range(depth + 1)]
        self.stream = torch.cuda.Stream(device=device)
        self.consumed = [None] * (depth + 1)   # event per slot: its readers are done
        self.owner = [None] * (depth + 1)      # (lo, hi) uploaded into the slot
        self.pending = {}                      # (lo, hi) -> (slot, chunk, done event)
        self.cursor = 0

    def issue(self, lo, hi):
        key = (lo, hi)
        if key in self.pending:
            return
        slot = self.cursor
        self.cursor = (self.cursor + 1) % len(self.buffers)
        self.pending.pop(self.owner[slot], None)  # a prediction nobody asked for
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])
                self.consumed[slot] = None
            chunk = self.buffers[slot][:hi - lo]
            chunk.copy_(_host_slice(self.data, lo, hi, chunk.dtype), non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.owner[slot] = key
        self.pending[key] = (slot, chunk, done)

    def take(self, lo, hi):
        self.issue(lo, hi)
        slot, chunk, done = self.pending.pop((lo, hi))
        torch.cuda.current_stream(self.device).wait_event(done)
        return slot, chunk

    def release(self, slot):
        used = torch.cuda.Event()
        used.record(torch.cuda.current_stream(self.device))
        self.consumed[slot] = used


_RINGS = {}  # id(data) -> (weakref to data, key, _HostRing)


def _ring_for(data, rows, device, depth):
    """The upload ring of ``data``, created on first use and dropped with the
    array (or by ``release_host_rings``)."""
    key = (rows, str(device), depth, tuple(data.shape[1:]), str(_staged_dtype(data)))
    hit = _RINGS.get(id(data))
    if hit is not None and hit[0]() is data and hit[1] == key:
        return hit[2]
    ring = _HostRing(data, rows, device, depth)
    try:
        ref = _weakref.ref(data, lambda _r, i=id(data): _RINGS.pop(i, None))
    except TypeError:  # not weak-referenceable: no reuse across epochs
        return ring
    _RINGS[id(data)] = (ref, key, ring)
    return ring


def release_host_rings():
    """Free every upload ring (Reconstruction.__exit__)."""
    _RINGS.clear()


def draw_sequence(num_batch, compact, comm=None):
    """The batch order of this epoch (rpie.py:95-98, lstsq.py:88-91) and a
    PREDICTION of the next epoch's: the generator is peeked and put back, so
    the draws are exactly the reference's; if something else draws from it in
    between, the prediction is wrong and only costs a wasted prefetch."""
    rng = tb_random.randomizer_np
    if compact:
        sequence = nxt = list(range(num_batch))
    else:
        sequence = [int(n) for n in rng.permutation(num_batch)]
        state = rng.bit_generator.state
        nxt = [int(n) for n in rng.permutation(num_batch)]
        rng.bit_generator.state = state
    if comm is not None and comm.size > 1:
        # every rank must visit the batches in the same order
        sequence, nxt = comm.bcast_object((sequence, nxt))
    return sequence, nxt


def peek_sequence(num_batch, compact, comm, predicted):
    """Prediction of the next epoch's batch order made at the END of an epoch.
    One process: the generator is peeked again, so draws made since the start of
    the epoch (the RANSAC subsets of the position fit) are accounted for.
    Several ranks: the prediction rank 0 broadcast with this epoch's order."""
    if comm is not None and comm.size > 1:
        return predicted
    if compact:
        return list(range(num_batch))
    rng = tb_random.randomizer_np
    state = rng.bit_generator.state
    nxt = [int(n) for n in rng.permutation(num_batch)]
    rng.bit_generator.state = state
    return nxt


class BatchStager:
    """Delivers the diffraction patterns of each batch as device tensors.

    Device-resident data is sliced.  Host (pinned) data is uploaded on a side
    stream piece by piece, ``depth`` pieces ahead of the compute stream, across
    batch boundaries and -- through a ring that outlives the solver call and
    the predicted next batch order -- across epoch boundaries, so the H2D copy
    of piece j+1 overlaps the kernels of piece j (the reference triple-buffers
    64-pattern chunks the same way, stream.py:359-404)."""

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
        self.depth = max(1, int(_os.environ.get('TB_STAGE_DEPTH', depth)))
        if chunk_positions is None and not self.resident:
            # pieces of about 256 MiB of float32 patterns (4096 at 128 x 128): few
            # enough launches per batch that their ramp-up does not show, small
            # enough that a piece is on the device well before its kernels.
            # Measured at BASELINE config 2 (uint16 stream, 5 batches of 20 000):
            # 4096 -> 108.7 ms per epoch, 8192 -> 109.6, whole batches -> 112.9
            pixels = int(np.prod(data.shape[1:]))
            default = max(256, min(4096, (256 << 20) // max(4 * pixels, 1)))
            chunk_positions = int(_os.environ.get('TB_STAGE_CHUNK', default))
        # flat list of (k, lo, hi) over the whole epoch
        self._chunk_positions, self._cuts = chunk_positions, cuts
        self._plan, self._first = self._make_plan(self.sequence, chunk_positions, cuts)
        self._ring = None
        if not self.resident:
            rows = max((hi - lo for _, lo, hi in self._plan), default=0)
            self._ring = _ring_for(data, rows, device, self.depth)
            for j in range(min(self.depth, len(self._plan))):
                self._issue(j)

    def prefetch_next(self, next_sequence):
        """Start the uploads of the first pieces of the next epoch (its
        predicted batch order); called by the solver once every batch of this
        epoch has been enqueued."""
        if self.resident or next_sequence is None:
            return
        plan = self._make_plan(list(next_sequence), self._chunk_positions, self._cuts)[0]
        for _, lo, hi in plan[:self.depth]:
            if hi > lo:
                self._ring.issue(lo, hi)

    def _make_plan(self, sequence, chunk_positions, cuts):
        plan, first = [], {}
        for k in range(len(sequence)):
            b = self.batches[sequence[k]]
            lo, hi = int(b[0]), int(b[-1]) + 1
            first[k] = len(plan)
            step = (hi - lo) if (self.resident or not chunk_positions or chunk_positions <= 0) \
                else chunk_positions
            cut = int(cuts[sequence[k]]) if cuts is not None else lo
            c = lo
            while c < hi or (c == lo and hi == lo):
                end = min(hi, c + max(step, 1))
                if c < cut < end:
                    end = cut
                plan.append((k, c, end))
                c = end
                if hi == lo:
                    break
        return plan, first

    def _range(self, k):
        b = self.batches[self.sequence[k]]
        return int(b[0]), int(b[-1]) + 1

    def _issue(self, j):
        """Start the upload of piece j of this epoch."""
        if j < len(self._plan):
            _, lo, hi = self._plan[j]
            if hi > lo:
                self._ring.issue(lo, hi)

    def chunks(self, k):
        """Yield ``(lo, hi, patterns)`` covering the k-th batch of the sequence.
        A yielded piece is valid until the next one is requested."""
        j = self._first[k]
        while j < len(self._plan) and self._plan[j][0] == k:
            _, lo, hi = self._plan[j]
            if self.resident:
                yield lo, hi, stage_data(self.data, lo, hi, self.device)
            elif hi == lo:
                yield lo, hi, self._ring.buffers[0][:0]
            else:
                slot, chunk = self._ring.take(lo, hi)
                for ahead in range(1, self.depth):
                    self._issue(j + ahead)
                yield lo, hi, chunk
                self._ring.release(slot)
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
