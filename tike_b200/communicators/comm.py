"""torch.distributed-backed communicator."""
from __future__ import annotations

import typing

import numpy as np
import torch
import torch.distributed as dist


class RowPlan:
    """Which object rows every rank touches and which it owns.

    Scan positions are split into row stripes (cluster.stripes_equal_count), so
    rank r only ever reads or writes the object rows ``touched[r] = [lo, hi)``
    that lie under its footprints.  ``own[r] = [bounds[r], bounds[r + 1])`` is
    a partition of all rows (rank r owns the rows from the top of its own
    footprints to the top of the next rank's).  A sum over ranks of an
    object-sized array (the reference's per-GPU ``psi_update_numerator``,
    preconditioner, ...) is then only needed where touched ranges overlap, and
    only by the ranks that touch those rows -- a halo exchange of about one
    probe height of rows between neighbours instead of an all-reduce of the
    whole object (replaces the Allreduce the reference sketches at
    _preconditioner.py:185, 201 and the per-epoch halo blend pool.py:415-476).
    """

    def __init__(self, touched, height: int):
        self.height = int(height)
        self.size = len(touched)
        self.touched = [(max(0, int(lo)), min(self.height, int(hi))) for lo, hi in touched]
        bounds = [0]
        for r in range(1, self.size):
            bounds.append(min(self.height, max(bounds[-1], self.touched[r][0])))
        bounds.append(self.height)
        self.bounds = bounds

    @classmethod
    def from_scan_rows(cls, row_minmax, probe_width: int, height: int):
        """``row_minmax[r] = (min, max)`` of rank r's scan row coordinates (or
        None for a rank without positions); a footprint covers the rows
        floor(y) ... floor(y) + probe_width (bilinear neighbour included)."""
        touched = [None] * len(row_minmax)
        nxt = int(height)  # a rank without positions owns no rows
        for r in range(len(row_minmax) - 1, -1, -1):
            mm = row_minmax[r]
            if mm is None:
                touched[r] = (nxt, nxt)
                continue
            lo = int(np.floor(mm[0]))
            touched[r] = (lo, int(np.floor(mm[1])) + int(probe_width) + 1)
            nxt = max(0, min(lo, int(height)))
        return cls(touched, height)

    def own(self, r):
        return self.bounds[r], self.bounds[r + 1]

    @staticmethod
    def _cut(a, b):
        lo, hi = max(a[0], b[0]), min(a[1], b[1])
        return (lo, hi) if hi > lo else None

    def to_owner(self, me):
        """[(peer, send rows, recv rows)]: rows I touch but ``peer`` owns, and
        rows ``peer`` touches but I own (either may be None)."""
        out = []
        for q in range(self.size):
            if q == me:
                continue
            send = self._cut(self.touched[me], self.own(q))
            recv = self._cut(self.touched[q], self.own(me))
            if send or recv:
                out.append((q, send, recv))
        return out

    def pairwise(self):
        """True when no object row is touched by more than two ranks (stripes
        taller than a footprint): the sum over ranks of a shared row is then a
        sum of two terms, which is the same on both sides whatever the order."""
        events = []
        for lo, hi in self.touched:
            if hi > lo:
                events += [(lo, 1), (hi, -1)]
        depth = 0
        for _, d in sorted(events):
            depth += d
            if depth > 2:
                return False
        return True

    def overlaps(self, me):
        """[(peer, rows)]: rows both ``me`` and ``peer`` touch."""
        out = []
        for q in range(self.size):
            if q != me:
                c = self._cut(self.touched[me], self.touched[q])
                if c:
                    out.append((q, c))
        return out

    def active(self, me):
        """Smallest row range holding everything rank ``me`` reads, writes or
        receives during an epoch: its touched rows and the rows of other
        ranks' contributions it sums as their owner."""
        lo, hi = self.touched[me]
        for _, _, recv in self.to_owner(me):
            if recv:
                lo, hi = (recv[0], recv[1]) if hi <= lo else (min(lo, recv[0]), max(hi, recv[1]))
        return lo, hi

    def shared_rows(self, me):
        """Rows of touched[me] that another rank touches as well."""
        out = []
        for q in range(self.size):
            if q != me:
                c = self._cut(self.touched[me], self.touched[q])
                if c:
                    out.append(c)
        return out


class Comm:
    """Collectives used by the reconstruction driver.

    ``size == 1`` (no process group) makes every method a no-op / identity so
    single-GPU runs need no rendezvous.  Reference call sites:
    ptycho.py:474-515 (probe mean, psi halo swap, transform mean, cost
    gather), _preconditioner.py:185,201 (commented-out Allreduce) and
    ptycho.py:946 (init rescale reduce).
    """

    def __init__(self, group=None, single: bool = False):
        """``single=True`` gives a one-rank communicator inside a multi-rank
        job (e.g. a reference run on one rank while the others wait)."""
        self.group = group
        self.plan = None        # RowPlan of the 'halo' data plane (ptycho.Reconstruction)
        self.batch_cuts = None  # per batch: index after the last boundary position
        if single:
            self.size, self.rank = 1, 0
        elif dist.is_available() and dist.is_initialized():
            self.size = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.size = 1
            self.rank = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # -- in-place tensor collectives ------------------------------------
    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            buf = torch.view_as_real(t) if t.is_complex() else t
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_mean_(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            self.allreduce_sum_(t)
            t /= self.size
        return t

    def allreduce_max_(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    # -- row-bounded reductions of object-sized arrays --------------------
    def _exchange(self, ops_spec):
        """ops_spec: [(peer, send tensor | None, recv tensor | None)] with
        contiguous tensors; one grouped send/recv."""
        ops = []
        for peer, send, recv in ops_spec:
            if send is not None:
                ops.append(dist.P2POp(dist.isend, torch.view_as_real(send) if send.is_complex()
                                      else send, peer, self.group))
            if recv is not None:
                ops.append(dist.P2POp(dist.irecv, torch.view_as_real(recv) if recv.is_complex()
                                      else recv, peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def halo_sum_(self, t: torch.Tensor, plan: RowPlan) -> torch.Tensor:
        """Make ``t`` (..., H, W) equal to the sum over all ranks on the rows
        this rank touches (``plan.touched[rank]``), in place.  Other rows are
        left as they are (nobody on this rank reads them).  When no row has
        more than two contributors (``plan.pairwise()``) one grouped send/recv
        with every overlapping neighbour does it; otherwise two rounds:
        contributions go to the owner of each row, who returns the complete
        sum to the ranks that touch the row (one summation order, so the
        replicas stay bit-identical)."""
        if self.size == 1:
            return t

        def rows(r):
            return t[..., r[0]:r[1], :]

        if plan.pairwise():
            # every shared row has exactly two contributors: one symmetric
            # exchange, both sides add (a + b == b + a bit for bit)
            spec, incoming = [], []
            for peer, r in plan.overlaps(self.rank):
                sbuf = rows(r).contiguous()
                rbuf = torch.empty_like(sbuf)
                spec.append((peer, sbuf, rbuf))
                incoming.append((r, rbuf))
            self._exchange(spec)
            for r, buf in incoming:
                rows(r).add_(buf)
            return t

        pairs = plan.to_owner(self.rank)
        if not pairs:
            return t

        # round 1: my contributions on rows owned by others -> their owner
        spec, incoming = [], []
        for peer, send, recv in pairs:
            sbuf = rows(send).contiguous() if send else None
            rbuf = torch.empty_like(rows(recv), memory_format=torch.contiguous_format) if recv else None
            spec.append((peer, sbuf, rbuf))
            if recv:
                incoming.append((recv, rbuf))
        self._exchange(spec)
        for r, buf in incoming:
            rows(r).add_(buf)
        # round 2: complete sums of my rows -> the ranks that touch them
        spec, incoming = [], []
        for peer, send, recv in pairs:
            sbuf = rows(recv).contiguous() if recv else None
            rbuf = torch.empty_like(rows(send), memory_format=torch.contiguous_format) if send else None
            spec.append((peer, sbuf, rbuf))
            if send:
                incoming.append((send, rbuf))
        self._exchange(spec)
        for r, buf in incoming:
            rows(r).copy_(buf)
        return t

    def halo_refresh_(self, t: torch.Tensor, plan: RowPlan) -> torch.Tensor:
        """Owner -> reader copy of the rows this rank touches but another rank
        owns (the second round of halo_sum_ alone): after an owner-side update
        every rank again holds current values on all the rows it reads."""
        if self.size == 1:
            return t
        pairs = plan.to_owner(self.rank)
        spec, incoming = [], []
        for peer, mine_at_peer, peers_at_me in pairs:
            sbuf = t[..., peers_at_me[0]:peers_at_me[1], :].contiguous() if peers_at_me else None
            rbuf = None
            if mine_at_peer:
                rbuf = torch.empty_like(t[..., mine_at_peer[0]:mine_at_peer[1], :],
                                        memory_format=torch.contiguous_format)
                incoming.append((mine_at_peer, rbuf))
            spec.append((peer, sbuf, rbuf))
        self._exchange(spec)
        for r, buf in incoming:
            t[..., r[0]:r[1], :].copy_(buf)
        return t

    def gather_owned_rows_(self, t: torch.Tensor, plan: RowPlan) -> torch.Tensor:
        """Every rank ends up with the owner's copy of every row of ``t``
        (..., H, W): one broadcast per owner (row counts differ per rank)."""
        if self.size == 1:
            return t
        for r in range(self.size):
            lo, hi = plan.own(r)
            if hi <= lo:
                continue
            view = t[..., lo:hi, :]
            if view.is_contiguous():
                self.bcast_(view, src=r)
            else:
                buf = view.contiguous()
                self.bcast_(buf, src=r)
                if r != self.rank:
                    view.copy_(buf)
        return t

    def bcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.size > 1:
            buf = torch.view_as_real(t) if t.is_complex() else t
            dist.broadcast(buf, src=src, group=self.group)
        return t

    def barrier(self):
        if self.size > 1:
            dist.barrier(group=self.group)

    # -- small host objects ----------------------------------------------
    def allgather_object(self, obj) -> list:
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def bcast_object(self, obj, src: int = 0):
        if self.size == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def reduce_cpu_sum(self, x: np.ndarray) -> np.ndarray:
        """Sum of small host arrays over ranks (ptycho.py:946)."""
        if self.size == 1:
            return x
        return np.sum(self.allgather_object(np.asarray(x)), axis=0)

    # -- stripe mode (reference-faithful halo blend) ---------------------
    def swap_edges(self, psi: torch.Tensor, overlap: int,
                   edges: typing.Sequence[int]) -> torch.Tensor:
        """Blend the band [edge, edge + overlap) of neighbouring stripes with
        linear ramps (pool.py:415-476) — rank i exchanges with i+1."""
        if self.size == 1:
            return psi
        if overlap < 1:
            raise ValueError(
                f"Overlap for swap_edges cannot be less than 1: {overlap}")
        for i in range(self.size - 1):
            lo = edges[i + 1]
            hi = lo + overlap
            if self.rank not in (i, i + 1):
                continue
            mine = psi[..., lo:hi, :].contiguous()
            theirs = torch.empty_like(mine)
            peer = i + 1 if self.rank == i else i
            a, b = torch.view_as_real(mine), torch.view_as_real(theirs)
            ops = [dist.P2POp(dist.isend, a, peer, self.group),
                   dist.P2POp(dist.irecv, b, peer, self.group)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            lower, upper = (mine, theirs) if self.rank == i else (theirs, mine)
            psi[..., lo:hi, :] = swap_edges_pair(lower, upper, overlap,
                                                 is_lower=self.rank == i)
        return psi


def swap_edges_pair(lower, upper, overlap: int, is_lower: bool):
    """Blended band for the lower (i) or upper (i+1) stripe owner:
    lower gets rampd*lower + rampu*upper, upper gets the same blend
    (pool.py:446-475)."""
    ramp = torch.linspace(0.0, 1.0, overlap + 2, device=lower.device,
                          dtype=torch.float32)[1:-1][..., None]
    return (1.0 - ramp) * lower + ramp * upper


def stitch_stripes(parts, stripe_start, probe_width: int):
    """Host-side stitch of per-rank objects (ObjectOptions.join_psi,
    object.py:154-167): rank i owns the rows from the centre of its first
    probe footprint to the centre of the next rank's; rank 0 also keeps
    everything above, the last rank everything below."""
    out = parts[0]
    half = probe_width // 2
    rows = out.shape[1]
    cuts = [int(s) + half for s in stripe_start[1:len(parts)]] + [rows]
    for i in range(1, len(parts)):
        out[:, cuts[i - 1]:cuts[i], :] = parts[i][:, cuts[i - 1]:cuts[i], :]
    return out
